"""Self-play steps on one engine (no replicas): moves/s for a board size / game count.
Usage: probe_selfplay.py [n=6] [games=5328] [rollouts=800] [steps=3]"""
import sys
import time

sys.path.insert(0, '.')
import tak_b200 as tb  # noqa: E402
from tak_b200 import weights as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
G = int(sys.argv[2]) if len(sys.argv) > 2 else 5328
R = int(sys.argv[3]) if len(sys.argv) > 3 else 800
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = tb.Engine(n, G, nodes_per_game=1 << 17, max_batch=G)
eng.net_create(n)
eng.net_load_weights(W.random_weights(n, seed=0))
eng.reset(0, G, 4)
print("forward", eng.net_forward_profile(0, G, 5))
eng.selfplay_begin(rollouts=R, half_komi=4, instant_win=1, exploit_ply=40, noise_ply=0, seed=1)
for i in range(steps):
    t = time.time()
    st = eng.selfplay_step(1)
    dt = time.time() - t
    print({"n": n, "games": G, "rollouts": R, "step": i, "device_ms": st.device_ms, "wall_s": dt,
           "moves_per_s": G / (st.device_ms / 1e3), "evals": st.evals, "launches": st.kernel_launches})
eng.close()
