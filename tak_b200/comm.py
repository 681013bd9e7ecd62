"""NCCL through the C ABI (include/taknative.h, "multi-GPU"): the communicator belongs to an `Engine` and every collective
runs on that engine's CUDA stream -- no torch.distributed involved.  One process per GPU; the 128-byte unique id is made
by one rank (`unique_id()`) and handed to the others by whatever channel the host program has (a pipe, a file, a TCP
store, MPI ...), exactly what the Rust `train` binary would do (train/src/main.rs:101-105,120 is the publish step this
serves; the reference itself is single-process).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ReplayRecord, check

ID_BYTES = 128


def unique_id() -> bytes:
    """ncclGetUniqueId (call on ONE rank, ship the bytes to the others)."""
    buf = (C.c_uint8 * ID_BYTES)()
    check(_lib.load().tak_comm_unique_id(buf, ID_BYTES))
    return bytes(buf)


class Comm:
    """The engine's communicator: `Comm(engine, uid, rank, world)` is collective over all ranks."""

    def __init__(self, engine, uid: bytes, rank: int, world: int):
        assert len(uid) == ID_BYTES
        self.engine, self.rank, self.world = engine, rank, world
        self.lib = engine.lib
        buf = (C.c_uint8 * ID_BYTES).from_buffer_copy(uid)
        check(self.lib.tak_comm_init(engine._h, buf, rank, world))

    def close(self):
        if self.engine is not None and getattr(self.engine, "_h", None):
            check(self.lib.tak_comm_destroy(self.engine._h))
        self.engine = None

    def bytes_moved(self) -> int:
        out = C.c_uint64()
        check(self.lib.tak_comm_info(self.engine._h, None, None, C.byref(out)))
        return out.value

    def broadcast_weights(self, blob: Optional[np.ndarray], root: int = 0) -> None:
        """Rank `root` publishes its fp32 weight blob; every rank's engine ends with that network loaded."""
        elems = self.engine.net_weights_size()
        ptr = None
        if self.rank == root:
            blob = np.ascontiguousarray(blob, dtype=np.float32)
            assert blob.size == elems
            ptr = blob.ctypes.data_as(C.POINTER(C.c_float))
        check(self.lib.net_broadcast_weights(self.engine._h, ptr, elems, root))

    def gather_replay(self, records: Sequence[ReplayRecord], cap: Optional[int] = None) -> List[ReplayRecord]:
        """All ranks' replay records, concatenated in rank order, on every rank."""
        k = len(records)
        arr = (ReplayRecord * max(k, 1))(*records)
        cnt = C.c_int32()
        if cap is None:                                   # first ask for the total (cap 0 fails with the count set)
            r = self.lib.selfplay_gather_replay(self.engine._h, arr, k, None, 0, C.byref(cnt))
            if r == 0 and cnt.value == 0:
                return []
            cap = cnt.value
        out = (ReplayRecord * max(cap, 1))()
        check(self.lib.selfplay_gather_replay(self.engine._h, arr, k, out, cap, C.byref(cnt)))
        return [ReplayRecord.from_buffer_copy(out[i]) for i in range(cnt.value)]

    def allreduce_gradients(self) -> None:
        """Sum of the ranks' gradient accumulators, in place, ordered before the next train_step on the engine stream."""
        check(self.lib.net_train_allreduce(self.engine._h))

    def sum_u64(self, values) -> List[int]:
        vals = [int(v) for v in (values if hasattr(values, "__len__") else [values])]
        buf = (C.c_uint64 * len(vals))(*vals)
        check(self.lib.tak_comm_sum_u64(self.engine._h, buf, len(vals)))
        return list(buf)

    def max_f64(self, values) -> List[float]:
        vals = [float(v) for v in (values if hasattr(values, "__len__") else [values])]
        buf = (C.c_double * len(vals))(*vals)
        check(self.lib.tak_comm_max_f64(self.engine._h, buf, len(vals)))
        return list(buf)
