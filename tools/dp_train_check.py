"""Data-parallel training check (run under torchrun with N ranks, one GPU each): every rank runs train_inner on its OWN
chunk, the gradient blobs are summed with an NCCL all-reduce (parallel.allreduce_gradients), every rank takes the Adam
step -- and the resulting weights must equal the single-process result of accumulating all N chunks before the step
(the reference's semantics: gradients of chunks add, network.rs:84-95).  Prints one JSON line from rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tak_b200 as tb  # noqa: E402
from tak_b200 import parallel as par  # noqa: E402
from tak_b200 import weights as W  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
B = 512


def chunk(seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.rand((B, 92, 6, 6), generator=g) < 0.15).float()
    pi = torch.zeros((B, 9036))
    idx = torch.randint(0, 9036, (B, 40), generator=g)
    pi.scatter_(1, idx, torch.rand((B, 40), generator=g))
    pi /= pi.sum(1, keepdim=True)
    z = (torch.randint(0, 3, (B,), generator=g) - 1).float()
    return x.numpy(), pi.numpy(), z.numpy()


blob = W.random_weights(6, seed=2)
eng = tb.Engine(6, 8, device=local, nodes_per_game=1 << 10, max_batch=8)
eng.net_create(6)
eng.net_load_weights(blob)
eng.train_begin(B)
loss = eng.train_chunk(*chunk(100 + rank))
par.allreduce_gradients(eng.train_grad_tensor())
torch.cuda.synchronize()
grads_dp = eng.train_get(1)
eng.train_step(1e-3, 1e-4)
w_dp = eng.train_get(0)
# every rank must hold the same weights after the step (BatchNorm running statistics are per rank: excluded)
names = [n for n, _ in W.spec(6)]
mask = np.concatenate([np.full(int(np.prod(s)), "running_" not in n) for n, s in W.spec(6)])
ref = torch.from_numpy(w_dp[mask]).to(dev)
lo, hi = ref.clone(), ref.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
spread = float((hi - lo).abs().max())
out = None
if rank == 0:
    single = tb.Engine(6, 8, device=local, nodes_per_game=1 << 10, max_batch=8)
    single.net_create(6)
    single.net_load_weights(blob)
    single.train_begin(B)
    for r in range(world):
        single.train_chunk(*chunk(100 + r))
    grads_1 = single.train_get(1)
    single.train_step(1e-3, 1e-4)
    w_1 = single.train_get(0)
    gerr = float(np.linalg.norm(grads_dp[mask] - grads_1[mask]) / np.linalg.norm(grads_1[mask]))
    werr = float(np.abs(w_dp[mask] - w_1[mask]).max())
    out = {"world": world, "chunk_positions": B, "loss_rank0": loss, "weights_spread_over_ranks": spread,
           "grad_rel_l2_vs_single_process": gerr, "weights_max_abs_diff_vs_single_process": werr,
           "ok": bool(spread == 0.0 and gerr < 2e-3 and werr <= 2.1e-3)}
dist.barrier()
dist.destroy_process_group()
if out:
    print(json.dumps(out))
