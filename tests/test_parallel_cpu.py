"""world_size-2 gloo test of the multi-GPU host logic (tak_b200/parallel.py): weight broadcast, ragged replay
gather, game-id sharding.  On the GPU box the same functions run over NCCL (bench.py)."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tak_b200 as tb
from tak_b200 import parallel as par


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    elems = 1000
    blob = np.arange(elems, dtype=np.float32) * 0.5 if rank == 0 else None
    t = par.broadcast_weights(blob, elems, dev)
    ok_w = bool(torch.equal(t, torch.arange(elems, dtype=torch.float32) * 0.5))
    base = par.game_id_base(rank, 16)
    recs = []
    for i in range(3 + 2 * rank):  # ragged: 3 records on rank 0, 5 on rank 1
        r = tb.ReplayRecord()
        r.game_id = base + i
        r.game_serial = rank
        r.result = float(1 - 2 * rank)
        r.n_children = 2
        r.moves[0], r.moves[1] = 7, 9
        r.visits[0], r.visits[1] = 100 + i, 200 + i
        recs.append(r)
    allr = par.gather_replay(recs, tb.ReplayRecord, dev)
    ids = [r.game_id for r in allr]
    ok_r = ids == [0, 1, 2, 16, 17, 18, 19, 20] and [r.visits[0] for r in allr][3] == 100
    g = torch.full((1000,), float(rank + 1))
    g[rank] += 10.0
    par.allreduce_gradients(g)          # in place: both ranks end with the sum
    want = torch.full((1000,), 3.0)
    want[0] += 10.0
    want[1] += 10.0
    ok_r = ok_r and bool(torch.equal(g, want))
    mx = par.max_over_ranks(float(rank + 1), dev)
    sm = par.sum_over_ranks(float(rank + 1), dev)
    out_q.put((rank, ok_w, ok_r, mx, sm))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_w, ok_r, mx, sm in res:
        assert ok_w and ok_r and mx == 2.0 and sm == 3.0
