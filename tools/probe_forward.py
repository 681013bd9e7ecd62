"""Small driver for ncu: a few Net6 forwards of G boards (default 4096) through the C ABI."""
import sys
sys.path.insert(0, '.')
import tak_b200 as tb
from tak_b200 import weights as W
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = tb.Engine(6, G, nodes_per_game=64, max_batch=G)
eng.net_create(6)
eng.net_load_weights(W.random_weights(6, seed=0))
eng.reset(0, G, 4)
print(eng.net_forward_profile(0, G, reps))
eng.close()
