"""The `train` loop on N GPUs at toy sizes (run under torchrun): sharded self-play -> replay all-gather -> data-parallel
training -> sharded pit -> next turn.  Prints one JSON line from rank 0 (evidence for DESIGN.md section 5)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tak_b200 as tb  # noqa: E402
from tak_b200 import train_loop as TL  # noqa: E402
from tak_b200 import weights as W  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
G = 256
blob = W.random_weights(6, seed=5)
cur = tb.Engine(6, G, device=local, nodes_per_game=1 << 13, max_batch=G)
cand = tb.Engine(6, G, device=local, nodes_per_game=1 << 13, max_batch=G)
for e in (cur, cand):
    e.net_create(6)
    e.net_load_weights(blob)
sp = dict(rollouts=16, half_komi=4, instant_win=1, exploit_ply=6, noise_ply=8, max_plies=30, seed=3)
quiet = lambda *_: None
t0 = time.perf_counter()
blob1, examples, res0 = TL.distributed_iteration(cur, cand, blob, [], 1, dev, min_new_examples=1600, selfplay_kw=sp, log=quiet)
n1 = len(examples)
blob2, examples, res1 = TL.distributed_iteration(cur, cand, blob1, examples, 2, dev, pit_games=8, pit_rollouts=4, pit_batch=8,
                                                 min_new_examples=400, selfplay_kw=sp, log=quiet,
                                                 train_kw=dict(chunk_size=64, chunks_in_step=4, lr=1e-3))
dt = time.perf_counter() - t0
# every rank must agree on everything
h = torch.tensor([float(np.abs(blob2).sum()), float(len(examples)), float(res1.wins), float(res1.losses)], dtype=torch.float64, device=dev)
lo, hi = h.clone(), h.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "examples_after_turn_1": n1, "examples_after_turn_2": len(examples),
                      "pit": [res1.wins, res1.losses, res1.draws], "accepted": bool(res1.win_rate() > TL.WIN_RATE_THRESHOLD),
                      "weights_changed": bool(not np.array_equal(blob2, blob1)), "ranks_agree": bool(torch.equal(lo, hi)),
                      "local_losses_first_last": [TL.train_network.last_losses[0], TL.train_network.last_losses[-1]],
                      "seconds": dt}))
dist.barrier()
dist.destroy_process_group()
