import sys, time; sys.path.insert(0,'.')
import numpy as np, tak_b200 as tb
from tak_b200 import weights as W
for G in (4096, 8192):
    eng = tb.Engine(6, G, nodes_per_game=1<<18, max_batch=G)
    eng.net_create(6); eng.net_load_weights(W.random_weights(6, seed=0))
    eng.reset(0, G, 4)
    print(G, eng.net_forward_profile(0, G, 5))
    eng.selfplay_begin(rollouts=800, half_komi=4, instant_win=1, exploit_ply=40, noise_ply=0, seed=1)
    for i in range(3):
        t=time.time(); st = eng.selfplay_step(1); dt=time.time()-t
        print(G, 'step', i, 'device_ms', st.device_ms, 'wall', dt, 'moves/s', G/ (st.device_ms/1e3), 'evals', st.evals, 'launches', st.kernel_launches)
    eng.close()
