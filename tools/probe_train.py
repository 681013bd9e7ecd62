"""Time Network::train_inner on the device: `boards` positions per chunk (reference: CHUNK_SIZE 500 examples x 8
symmetries = 4000, network.rs:19), synthetic inputs resident in HBM.  Usage: probe_train.py [boards=4000] [reps=5] [arch=6]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tak_b200 as tb  # noqa: E402
from tak_b200 import weights as W  # noqa: E402

boards = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
arch = int(sys.argv[3]) if len(sys.argv) > 3 else 6
C_IN, P = tb.input_channels(arch), tb.policy_size(arch)
eng = tb.Engine(arch, 8, nodes_per_game=1 << 10, max_batch=8)
eng.net_create(arch)
eng.net_load_weights(W.random_weights(arch, seed=0))
eng.train_begin(boards)
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.rand((boards, C_IN, arch, arch), device="cuda", generator=g) < 0.15).float()
pi = torch.zeros((boards, P), device="cuda")
idx = torch.randint(0, P, (boards, 60), device="cuda", generator=g)
pi.scatter_(1, idx, torch.rand((boards, 60), device="cuda", generator=g))
pi /= pi.sum(1, keepdim=True)
z = (torch.randint(0, 3, (boards,), device="cuda", generator=g) - 1).float()
torch.cuda.synchronize()
for _ in range(2):
    loss = eng.train_chunk(x, pi, z)
ms = []
for _ in range(reps):
    loss = eng.train_chunk(x, pi, z)
    ms.append(eng.train_stats()["ms_last_chunk"])
eng.train_step()
flop = 3 * (368_197_632 if arch == 6 else 132_198_400) * boards
print({"arch": arch, "boards": boards, "ms_per_chunk": float(np.mean(ms)), "positions_per_s": boards / (np.mean(ms) * 1e-3),
       "tflops_fwd_bwd": flop / (np.mean(ms) * 1e-3) / 1e12, "loss": loss})
eng.train_end()
