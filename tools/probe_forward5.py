"""Net5 forward breakdown: tower vs heads (5x5 is configs[1], a parity config; this shows where its time goes)."""
import sys
sys.path.insert(0, '.')
import tak_b200 as tb
from tak_b200 import weights as W
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = tb.Engine(5, G, nodes_per_game=64, max_batch=G)
eng.net_create(5)
eng.net_load_weights(W.random_weights(5, seed=0))
eng.reset(0, G, 4)
p = eng.net_forward_profile(0, G, 5)
print({"boards": G, **p, "head_ms": p["ms_forward"] - p["ms_conv"], "evals_per_s": G / (p["ms_forward"] * 1e-3)})
eng.close()
