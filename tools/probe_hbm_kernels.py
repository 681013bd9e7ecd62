"""Driver for the ncu captures of the HBM-bound kernels around the conv tower (profiles/r02_hbm_kernels.md):
6x6 self-play in the bench's configuration (instant-win on, Dirichlet noise, sampled picks, 800 rollouts) on 888 games
(one conv tile per SM x 6 boards; the per-game traffic of the search kernels does not depend on the number of games), then
one Network::policy_eval of 888 positions (k_encode / k_value / k_policy_stats_conv) and one 8x8 playout batch.
  python tools/probe_hbm_kernels.py [plies=3] [rollouts=800]"""
import sys
sys.path.insert(0, '.')
import numpy as np  # noqa: E402
import tak_b200 as tb  # noqa: E402
from tak_b200 import weights as W  # noqa: E402

plies = int(sys.argv[1]) if len(sys.argv) > 1 else 3
R = int(sys.argv[2]) if len(sys.argv) > 2 else 800
G = 888
eng = tb.Engine(6, G, nodes_per_game=1 << 17, max_batch=G)
eng.net_create(6)
eng.net_load_weights(W.random_weights(6, seed=0))
eng.selfplay_begin(rollouts=R, half_komi=4, instant_win=1, exploit_ply=40, noise_ply=80, noise_alpha=0.2, noise_ratio=0.3,
                   seed=0x7A4B)
for i in range(plies):
    st = eng.selfplay_step(1)
    print({"ply": i, "device_ms": st.device_ms, "evals": st.evals, "rollouts": st.rollouts, "launches": st.kernel_launches})
states = eng.download(np.arange(G, dtype=np.int32))
pol, val = eng.policy_eval(states)
print("policy_eval", pol.shape, float(val.mean()))
eng.close()
eng = tb.Engine(8, 148 * 64, nodes_per_game=64)
eng.reset(0, 148 * 64, 4)
print("playouts", eng.playouts(0, 148 * 64, 7, 100_000, 0)[2])
eng.close()
