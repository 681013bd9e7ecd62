"""Multi-GPU plumbing of the self-play path (one process per GPU).

Games never interact, so each rank owns a contiguous block of global game ids, its own search trees and a network
replica; the rollout loop has no collective.  Exactly two exchanges exist:
  * `broadcast_weights`  rank 0's fp32 weight blob -> every rank
  * `gather_replay`      fixed-size replay records of all ranks -> every rank
The training step (next row N1) adds the one collective data-parallel training needs:
  * `allreduce_gradients` sum of every rank's fp32 gradient blob, in place, before the Adam step -- every rank then takes
                          the identical step, which is the reference's single-process step over world x as many chunks
On GPUs these run INSIDE the C ABI (tak_b200.comm.Comm over csrc/comm.cu: NCCL on the engine's stream, ordered with the
kernels around it); this module is the thin caller.  The torch.distributed versions below are the host-side fallback
used by the CPU tests (gloo, world_size 2) and by callers that have a process group but no engine communicator.
The reference has no distributed code (single process, single GPU: alpha-tak/src/lib.rs:21-23).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def game_id_base(rank: int, games_per_rank: int) -> int:
    """Global id of local game 0: ranks own [rank*G, (rank+1)*G)."""
    return rank * games_per_rank


def broadcast_weights(blob: np.ndarray | None, elems: int, device: torch.device, src: int = 0) -> torch.Tensor:
    """Returns the fp32 blob as a tensor on `device` on every rank (only `src` needs to pass `blob`)."""
    t = torch.empty(elems, dtype=torch.float32, device=device)
    if not dist.is_initialized() or dist.get_rank() == src:
        assert blob is not None and blob.size == elems
        t.copy_(torch.from_numpy(np.ascontiguousarray(blob, dtype=np.float32)))
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def gather_replay(records: Sequence, record_type, device: torch.device) -> List:
    """all_gather of ragged lists of fixed-size ctypes records; returns the concatenation in rank order."""
    size = C.sizeof(record_type)
    raw = b"".join(bytes(r) for r in records)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(records)
    world = dist.get_world_size()
    cnt = torch.tensor([len(records)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    payload = torch.zeros(mx * size, dtype=torch.uint8, device=device)
    if raw:
        payload[: len(raw)].copy_(torch.from_numpy(np.frombuffer(raw, dtype=np.uint8).copy()))
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload)
    out = []
    for r, c in enumerate(counts):
        buf = gathered[r][: c * size].cpu().numpy().tobytes()
        out += [record_type.from_buffer_copy(buf[i * size:(i + 1) * size]) for i in range(c)]
    return out


def allreduce_gradients(grad: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of the gradient blob (`Engine.train_grad_tensor()`: a zero-copy view of the engine's
    accumulator).  Gradients of chunks ADD in the reference (`total_loss.backward()` per chunk, one `opt.step()` per
    CHUNKS_IN_STEP chunks, network.rs:84-95), so the sum over ranks is the single-process gradient of all their chunks.

    torch.distributed enqueues NCCL on ITS stream and returns; the engine's Adam kernel runs on the engine's own
    non-blocking stream, which nothing orders after it.  So this fallback waits for the reduction to finish before it
    returns.  (`Comm.allreduce_gradients` needs no wait: it issues NCCL on the engine's stream.)"""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if grad.is_cuda:
            torch.cuda.synchronize(grad.device)      # the chunks' kernels (engine stream) have written `grad`
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
        if grad.is_cuda:
            torch.cuda.synchronize(grad.device)      # ... and the sum is complete before net_train_step is enqueued
    return grad


def engine_allreduce(eng, comm=None) -> None:
    """The all-reduce `train_loop.train_network` calls before each step: through the engine's communicator when there is
    one (NCCL on the engine stream), else the torch.distributed fallback above."""
    if comm is not None:
        comm.allreduce_gradients()
    else:
        allreduce_gradients(eng.train_grad_tensor())


def max_over_ranks(x: float, device: torch.device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device: torch.device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
