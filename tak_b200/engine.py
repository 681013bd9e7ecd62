"""Host-side mirror of the reference's interface for the hot path, over the C ABI (include/taknative.h).

`Engine` is the batched device object (thousands of games / search trees per GPU); `Game`, `Node` and `Network`
mirror `tak::Game<N>`, `alpha_tak::Node` and `alpha_tak::Network::policy_eval` one-to-one so the parity tests read
like the reference's own tests (tak/tests/*.rs, alpha-tak/src/search/tests.rs).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import EngineConfig, ReplayRecord, SelfplayConfig, SelfplayStats, TakNativeError, TakState, check

RESULT_ONGOING, RESULT_WHITE, RESULT_BLACK, RESULT_DRAW, RESULT_FLAG = 0, 1, 2, 3, 0x10


def _ids(ids) -> Tuple[C.Array, int]:
    arr = np.ascontiguousarray(ids, dtype=np.int32)
    return arr.ctypes.data_as(C.POINTER(C.c_int32)), int(arr.size), arr


# ---- cold host helpers (takparse surface) -----------------------------------------------------------------
def parse_move(text: str, n: int) -> int:
    """takparse `Move::from_str` -> u16 move."""
    out = C.c_uint16()
    check(_lib.load().tak_ptn_parse(n, text.encode(), C.byref(out)))
    return out.value


def format_move(move: int, n: int) -> str:
    buf = C.create_string_buffer(32)
    check(_lib.load().tak_ptn_format(n, move, buf, 32))
    return buf.value.decode()


def move_index(move: int, n: int) -> int:
    """alpha_tak::search::move_index (move_map.rs:19-48)."""
    out = C.c_int32()
    check(_lib.load().tak_move_index(n, move, C.byref(out)))
    return out.value


def policy_size(n: int) -> int:
    out = C.c_int32()
    check(_lib.load().tak_policy_size(n, C.byref(out)))
    return out.value


def input_channels(n: int) -> int:
    out = C.c_int32()
    check(_lib.load().net_input_channels(n, C.byref(out)))
    return out.value


def boards_per_tile(n: int) -> int:
    """Positions per 256-row tile of the conv tower (sizing hint: 148 SMs x k tiles x this many games)."""
    out = C.c_int32()
    check(_lib.load().net_boards_per_tile(n, C.byref(out)))
    return out.value


def state_init(n: int, half_komi: int = 0) -> TakState:
    s = TakState()
    check(_lib.load().tak_state_init(n, half_komi, C.byref(s)))
    return s


def tps_format(state: TakState) -> str:
    buf = C.create_string_buffer(8192)
    check(_lib.load().tak_tps_format(C.byref(state), buf, 8192))
    return buf.value.decode()


def tps_parse(n: int, text: str) -> TakState:
    s = TakState()
    check(_lib.load().tak_tps_parse(n, text.encode(), C.byref(s)))
    return s


class Engine:
    """Owns one GPU's game states, search trees and network (`tak_engine_t`)."""

    def __init__(self, n: int, max_games: int, device: int = 0, nodes_per_game: int = 0, max_batch: int = 0):
        self.lib = _lib.load()
        self.n = n
        self.max_games = max_games
        self.device = device
        cfg = EngineConfig(device=device, n=n, max_games=max_games, nodes_per_game=nodes_per_game,
                           max_batch=max_batch)
        h = C.c_void_p()
        check(self.lib.tak_engine_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.policy_size = policy_size(n)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.tak_engine_destroy(self._h)
            self._h = None
        for ptr in getattr(self, "_pinned", []):
            self.lib.tak_host_free(ptr)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.tak_engine_sync(self._h))

    # ---- tak::Game, batched -----------------------------------------------------------------------------
    def reset(self, first: int = 0, count: Optional[int] = None, half_komi: int = 0):
        check(self.lib.tak_games_reset(self._h, first, self.max_games - first if count is None else count, half_komi))

    def upload(self, ids, states: Sequence[TakState]):
        p, k, _keep = _ids(ids)
        arr = (TakState * k)(*states)
        check(self.lib.tak_games_upload(self._h, p, k, arr))

    def download(self, ids) -> List[TakState]:
        p, k, _keep = _ids(ids)
        arr = (TakState * k)()
        check(self.lib.tak_games_download(self._h, p, k, arr))
        return list(arr)

    def possible_moves(self, ids) -> List[np.ndarray]:
        """`Game::possible_moves` for each listed game, in the reference's order."""
        p, k, _keep = _ids(ids)
        cap = max(1024, 512 * k)
        while True:
            moves = np.zeros(cap, dtype=np.uint16)
            offs = np.zeros(k + 1, dtype=np.int32)
            r = self.lib.tak_possible_moves(self._h, p, k, moves.ctypes.data_as(C.POINTER(C.c_uint16)),
                                            offs.ctypes.data_as(C.POINTER(C.c_int32)), cap)
            if r == -34 and cap < (1 << 26):
                cap *= 4
                continue
            check(r)
            return [moves[offs[i]:offs[i + 1]].copy() for i in range(k)]

    def play(self, ids, moves) -> np.ndarray:
        """`Game::play`; returns per-game status (0 or the PlayError code)."""
        p, k, _keep = _ids(ids)
        mv = np.ascontiguousarray(moves, dtype=np.uint16)
        assert mv.size == k
        st = np.zeros(k, dtype=np.int32)
        check(self.lib.tak_play(self._h, p, mv.ctypes.data_as(C.POINTER(C.c_uint16)), k,
                                st.ctypes.data_as(C.POINTER(C.c_int32))))
        return st

    def result(self, ids) -> np.ndarray:
        p, k, _keep = _ids(ids)
        out = np.zeros(k, dtype=np.uint8)
        check(self.lib.tak_result(self._h, p, k, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def perft(self, root: TakState, depth: int) -> int:
        out = C.c_uint64()
        check(self.lib.tak_perft(self._h, C.byref(root), depth, C.byref(out)))
        return out.value

    def perft_multi(self, roots: Sequence[TakState], depth: int) -> int:
        """Sum of perft(root, depth) over `roots`, expanded together (one frontier)."""
        arr = (TakState * max(len(roots), 1))(*roots)
        out = C.c_uint64()
        check(self.lib.tak_perft_multi(self._h, arr, len(roots), depth, C.byref(out)))
        return out.value

    def frontier(self, root: TakState, depth: int):
        """(positions `depth` plies below `root` in move-generation order, number of lines that ended earlier).
        perft.rs:3-18 counts a finished game as 1 at whatever depth it ends, so
            perft(root, d) == ended + perft_multi(positions, d - depth).
        Host-driven through result / possible_moves / play on the engine's slots; meant for shallow depths."""
        level, ended = [root], 0
        for _ in range(depth):
            nxt: List[TakState] = []
            for lo in range(0, len(level), self.max_games):
                part = level[lo:lo + self.max_games]
                ids = list(range(len(part)))
                self.upload(ids, part)
                res = self.result(ids)
                lists = self.possible_moves(ids)
                parents, moves = [], []
                for st, r, mv in zip(part, res, lists):
                    if (int(r) & 3) != RESULT_ONGOING:
                        ended += 1
                        continue
                    parents += [st] * len(mv)
                    moves += [int(m) for m in mv]
                for lo2 in range(0, len(parents), self.max_games):
                    ids2 = list(range(min(self.max_games, len(parents) - lo2)))
                    self.upload(ids2, parents[lo2:lo2 + len(ids2)])
                    assert not self.play(ids2, moves[lo2:lo2 + len(ids2)]).any()
                    nxt += self.download(ids2)
            level = nxt
        return level, ended

    def perft_stats(self):
        ms, mat, launches = C.c_double(), C.c_uint64(), C.c_uint64()
        check(self.lib.tak_perft_stats(self._h, C.byref(ms), C.byref(mat), C.byref(launches)))
        return {"ms": ms.value, "materialised": mat.value, "launches": launches.value}

    def perft_profile(self):
        out = (C.c_double * 6)()
        check(self.lib.tak_perft_profile(self._h, out))
        return {"ms": out[0], "expand_ms": out[1], "materialised": int(out[2]), "launches": int(out[3]),
                "top_children": int(out[4]), "top_ms": out[5]}

    def playouts(self, first: int, count: int, seed: int, max_plies: int, ply_spread: int = 0, game_id_base: int = 0):
        """Uniform-random playouts of games [first, first+count) on the device (tak_playouts): returns
        (plies added per game, final GameResult per game, {"plies", "generated", "ms"})."""
        plies = np.zeros(max(count, 1), dtype=np.int32)
        res = np.zeros(max(count, 1), dtype=np.uint8)
        tot = (C.c_uint64 * 2)()
        ms = C.c_double()
        check(self.lib.tak_playouts(self._h, first, count, seed, game_id_base, max_plies, ply_spread,
                                    plies.ctypes.data_as(C.POINTER(C.c_int32)),
                                    res.ctypes.data_as(C.POINTER(C.c_uint8)), tot, C.byref(ms)))
        return plies[:count], res[:count], {"plies": int(tot[0]), "generated": int(tot[1]), "ms": ms.value}

    # ---- alpha_tak::Network --------------------------------------------------------------------------------
    def net_create(self, arch: int):
        check(self.lib.net_create(self._h, arch))

    def net_weights_size(self) -> int:
        out = C.c_int64()
        check(self.lib.net_weights_size(self._h, C.byref(out)))
        return out.value

    def net_load_weights(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        check(self.lib.net_load_weights(self._h, blob.ctypes.data_as(C.POINTER(C.c_float)), blob.size))

    def net_load_weights_device(self, ptr: int, elems: int):
        check(self.lib.net_load_weights_device(self._h, C.c_void_p(ptr), elems))

    def game_repr(self, states: Sequence[TakState]) -> np.ndarray:
        b = len(states)
        c = input_channels(self.n)
        out = np.zeros((b, c, self.n, self.n), dtype=np.float32)
        arr = (TakState * b)(*states)
        check(self.lib.net_game_repr(self._h, arr, b, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def pinned_array(self, shape, dtype=np.float32) -> np.ndarray:
        """A numpy array in page-locked host memory (tak_host_alloc); it lives as long as the engine object."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        ptr = C.c_void_p()
        check(self.lib.tak_host_alloc(max(1, n * dtype.itemsize), C.byref(ptr)))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_uint8 * (n * dtype.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def policy_eval(self, states, out: Optional[Tuple[np.ndarray, np.ndarray]] = None) -> Tuple[np.ndarray, np.ndarray]:
        """`Network::policy_eval(&[Game])` -> (policy [B, policy_size], eval [B]).  `states`: a sequence of TakState or a
        ctypes array of them; `out`: preallocated (policy, eval) arrays, e.g. from `pinned_array` (DMA instead of the
        driver's staged copy into pageable memory)."""
        b = len(states)
        if out is not None:
            pol, val = out
            assert pol.shape == (b, self.policy_size) and val.shape == (b,) and pol.dtype == val.dtype == np.float32
            assert pol.flags.c_contiguous and val.flags.c_contiguous
        else:
            pol = np.zeros((b, self.policy_size), dtype=np.float32)
            val = np.zeros(b, dtype=np.float32)
        if b == 0:
            return pol, val
        arr = states if isinstance(states, C.Array) else (TakState * b)(*states)
        check(self.lib.net_policy_eval(self._h, arr, b, pol.ctypes.data_as(C.POINTER(C.c_float)),
                                       val.ctypes.data_as(C.POINTER(C.c_float))))
        return pol, val

    def policy_logits(self, states: Sequence[TakState]) -> Tuple[np.ndarray, np.ndarray]:
        """The forward pass of `policy_eval`, returning the pre-softmax policy logits [B, policy_size] and eval [B]."""
        b = len(states)
        lg = np.zeros((b, self.policy_size), dtype=np.float32)
        val = np.zeros(b, dtype=np.float32)
        if b == 0:
            return lg, val
        arr = (TakState * b)(*states)
        check(self.lib.net_policy_logits(self._h, arr, b, lg.ctypes.data_as(C.POINTER(C.c_float)),
                                         val.ctypes.data_as(C.POINTER(C.c_float))))
        return lg, val

    def net_forward_timed(self, first: int, count: int, reps: int) -> float:
        ms = C.c_double()
        check(self.lib.net_forward_timed(self._h, first, count, reps, C.byref(ms)))
        return ms.value

    def net_forward_profile(self, first: int, count: int, reps: int):
        out = (C.c_double * 4)()
        check(self.lib.net_forward_profile(self._h, first, count, reps, out))
        return {"ms_forward": out[0], "ms_conv": out[1], "conv_launches": int(out[2]), "flop": out[3]}

    # ---- alpha_tak::Node, batched -------------------------------------------------------------------------------
    # ---- Network::train (network.rs:37-97), Net6 ----
    def train_begin(self, max_boards: int):
        check(self.lib.net_train_begin(self._h, max_boards))

    def train_chunk(self, inputs, pi, z):
        """train_inner for one chunk: numpy host arrays [b,C,n,n] / [b,P] / [b], or torch CUDA tensors of those shapes
        (passed by device pointer).  Returns (loss_p, loss_z); gradients accumulate until train_step."""
        on_device = hasattr(inputs, "data_ptr")
        if on_device:
            for t in (inputs, pi, z):
                assert t.is_cuda and t.is_contiguous() and t.dtype.is_floating_point and t.element_size() == 4
            b = int(inputs.shape[0])
            ptrs = [t.data_ptr() for t in (inputs, pi, z)]
        else:
            inputs = np.ascontiguousarray(inputs, dtype=np.float32)
            pi = np.ascontiguousarray(pi, dtype=np.float32)
            z = np.ascontiguousarray(z, dtype=np.float32)
            b = int(inputs.shape[0])
            ptrs = [a.ctypes.data for a in (inputs, pi, z)]
        assert pi.shape[0] == b and z.shape[0] == b
        out = (C.c_float * 2)()
        check(self.lib.net_train_chunk(self._h, ptrs[0], ptrs[1], ptrs[2], b, 1 if on_device else 0, out))
        return float(out[0]), float(out[1])

    def train_step(self, lr: float = 1e-4, weight_decay: float = 1e-4):       # network.rs:14-15
        check(self.lib.net_train_step(self._h, lr, weight_decay))

    def train_get(self, what: int = 0) -> np.ndarray:
        """0 weights (load them with net_load_weights to search with the trained network), 1 gradients, 2/3 Adam m/v."""
        out = np.zeros(self.net_weights_size(), dtype=np.float32)
        check(self.lib.net_train_get(self._h, what, out.ctypes.data_as(C.POINTER(C.c_float)), out.size))
        return out

    def train_grad_tensor(self):
        """The gradient blob as a zero-copy torch CUDA tensor (for torch.distributed.all_reduce in data-parallel runs)."""
        import torch
        ptr, n = C.c_void_p(), C.c_int64()
        check(self.lib.net_train_grad_ptr(self._h, C.byref(ptr), C.byref(n)))

        class _Mem:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": "<f4", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(_Mem(), device=torch.device("cuda", self.device))

    def train_stats(self):
        ms, chunks, steps = C.c_double(), C.c_int32(), C.c_int32()
        check(self.lib.net_train_stats(self._h, C.byref(ms), C.byref(chunks), C.byref(steps)))
        return {"ms_last_chunk": ms.value, "chunks_pending": chunks.value, "steps": steps.value}

    def train_end(self):
        check(self.lib.net_train_end(self._h))

    def tree_reset(self, ids):
        p, k, _keep = _ids(ids)
        check(self.lib.mcts_tree_reset(self._h, p, k))

    def virtual_rollout(self, ids, k_per_game: int = 1):
        p, k, _keep = _ids(ids)
        check(self.lib.mcts_virtual_rollout(self._h, p, k, k_per_game))

    def pending(self, with_states: bool = True):
        cnt = C.c_int32()
        check(self.lib.mcts_pending(self._h, C.byref(cnt), None, None, 0))
        k = cnt.value
        gids = np.zeros(max(k, 1), dtype=np.int32)
        states = (TakState * max(k, 1))()
        if k:
            check(self.lib.mcts_pending(self._h, C.byref(cnt), gids.ctypes.data_as(C.POINTER(C.c_int32)),
                                        states if with_states else None, k))
        return gids[:k].copy(), list(states)[:k]

    def devirtualize(self):
        check(self.lib.mcts_devirtualize(self._h))

    def reserve_pending(self, k: int):
        check(self.lib.mcts_reserve_pending(self._h, k))

    def devirtualize_first(self, ids, counts):
        """Back up only the oldest counts[i] queued leaves of game ids[i]; everything else stays queued."""
        p, k, _keep = _ids(ids)
        c = (C.c_int32 * max(k, 1))(*[int(x) for x in counts])
        check(self.lib.mcts_devirtualize_first(self._h, p, k, c))

    def devirtualize_with(self, policy: np.ndarray, value: np.ndarray):
        policy = np.ascontiguousarray(policy, dtype=np.float32)
        value = np.ascontiguousarray(value, dtype=np.float32)
        check(self.lib.mcts_devirtualize_with(self._h, policy.ctypes.data_as(C.POINTER(C.c_float)),
                                              value.ctypes.data_as(C.POINTER(C.c_float)), value.size))

    def player_rollouts(self, ids, batch: int, reps: int = 1):
        """Player::rollout x reps, fused: queue a new batch per game, then back up the batch already outstanding."""
        p, k, _keep = _ids(ids)
        check(self.lib.mcts_player_rollouts(self._h, p, k, batch, reps))

    def rollouts(self, ids, n_rollouts: int):
        p, k, _keep = _ids(ids)
        check(self.lib.mcts_rollouts(self._h, p, k, n_rollouts))

    def children(self, gid: int):
        cap = 4096
        mv = np.zeros(cap, dtype=np.uint16)
        vis = np.zeros(cap, dtype=np.uint32)
        pri = np.zeros(cap, dtype=np.float32)
        rew = np.zeros(cap, dtype=np.float32)
        cnt = C.c_int32()
        check(self.lib.mcts_children(self._h, gid, mv.ctypes.data_as(C.POINTER(C.c_uint16)),
                                     vis.ctypes.data_as(C.POINTER(C.c_uint32)),
                                     pri.ctypes.data_as(C.POINTER(C.c_float)),
                                     rew.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(cnt)))
        k = cnt.value
        return mv[:k].copy(), vis[:k].copy(), pri[:k].copy(), rew[:k].copy()

    def debug(self, gid: int, depth: int = 10):
        """Node::debug(depth) (search/debug.rs:9-24) of game `gid`'s root -> analysis.NodeDebugInfo."""
        from ._lib import MoveInfoRecord
        from .analysis import MoveInfo, NodeDebugInfo
        cap = 4096
        buf = (MoveInfoRecord * cap)()
        cnt = C.c_int32()
        check(self.lib.mcts_debug(self._h, gid, depth, buf, cap, C.byref(cnt)))
        return NodeDebugInfo([MoveInfo.from_record(buf[i], self.n) for i in range(cnt.value)])

    def children_batch(self, ids, stride: int = 256):
        """`Node::improved_policy` of many roots at once: (moves [n, stride], visits [n, stride], counts [n])."""
        p, k, _keep = _ids(ids)
        mv = np.zeros((k, stride), dtype=np.uint16)
        vis = np.zeros((k, stride), dtype=np.uint32)
        cnt = np.zeros(k, dtype=np.int32)
        check(self.lib.mcts_children_batch(self._h, p, k, mv.ctypes.data_as(C.POINTER(C.c_uint16)),
                                           vis.ctypes.data_as(C.POINTER(C.c_uint32)),
                                           cnt.ctypes.data_as(C.POINTER(C.c_int32)), stride))
        return mv, vis, cnt

    def root(self, gid: int):
        v, vv, r = C.c_uint32(), C.c_uint32(), C.c_float()
        check(self.lib.mcts_root(self._h, gid, C.byref(v), C.byref(vv), C.byref(r)))
        return v.value, vv.value, r.value

    def pick_move(self, ids) -> np.ndarray:
        p, k, _keep = _ids(ids)
        out = np.zeros(k, dtype=np.uint16)
        check(self.lib.mcts_pick_move(self._h, p, k, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def pick_move_sampled(self, ids, seed: int) -> np.ndarray:
        """`Node::pick_move(false)`: visit-weighted draw per listed game (counter-based: seed x game id)."""
        p, k, _keep = _ids(ids)
        out = np.zeros(k, dtype=np.uint16)
        check(self.lib.mcts_pick_move_sampled(self._h, p, k, seed, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def tree_play(self, ids, moves):
        p, k, _keep = _ids(ids)
        mv = np.ascontiguousarray(moves, dtype=np.uint16)
        check(self.lib.mcts_play(self._h, p, mv.ctypes.data_as(C.POINTER(C.c_uint16)), k))

    def apply_dirichlet(self, ids, alpha: float, ratio: float, seed: int):
        p, k, _keep = _ids(ids)
        check(self.lib.mcts_apply_dirichlet(self._h, p, k, alpha, ratio, seed))

    # ---- train::self_play_parallel ----------------------------------------------------------------------------
    def selfplay_begin(self, **kw):
        cfg = SelfplayConfig(rollouts=kw.get("rollouts", 800), half_komi=kw.get("half_komi", 4),
                             instant_win=kw.get("instant_win", 1), exploit_ply=kw.get("exploit_ply", 40),
                             noise_ply=kw.get("noise_ply", 0), noise_alpha=kw.get("noise_alpha", 0.2),
                             noise_ratio=kw.get("noise_ratio", 0.3), seed=kw.get("seed", 0x7A4B),
                             max_plies=kw.get("max_plies", 0), game_id_base=kw.get("game_id_base", 0))
        cfg.reserved[0] = 1 if kw.get("keep_positions", False) else 0     # start from the positions the slots hold now
        check(self.lib.selfplay_begin(self._h, C.byref(cfg)))

    def selfplay_step(self, moves: int = 1) -> SelfplayStats:
        st = SelfplayStats()
        check(self.lib.selfplay_step(self._h, moves, C.byref(st)))
        return st

    def selfplay_drain(self, cap: Optional[int] = None) -> List[ReplayRecord]:
        """Completed replay records (games that finished), at most `cap` of them (default: all that are waiting).  The
        records are copies: they do not keep the staging array alive."""
        cnt = C.c_int32()
        check(self.lib.selfplay_drain(self._h, None, 0, C.byref(cnt)))      # how many completed records are waiting
        k = cnt.value if cap is None else min(cap, cnt.value)
        if k <= 0:
            return []
        arr = (ReplayRecord * k)()
        check(self.lib.selfplay_drain(self._h, arr, k, C.byref(cnt)))
        return [ReplayRecord.from_buffer_copy(arr[i]) for i in range(cnt.value)]


    # ---- alpha_tak::Example::to_tensors, batched (example.rs:63-78) -------------------------------------------
    def examples_to_tensors(self, records: Sequence[ReplayRecord], on_device: bool = False):
        """8-fold symmetry augmentation of replay records on the device:
        (inputs [8k, C, n, n], pi [8k, policy_size], z [8k]), row 8e+s = symmetry s of example e -- as host arrays, or
        with on_device as torch CUDA tensors left in HBM (what train_chunk consumes without a host round trip)."""
        k = len(records)
        arr = (ReplayRecord * max(k, 1))(*records)
        c = input_channels(self.n)
        if on_device:
            import torch
            dev = torch.device("cuda", self.device)
            inputs = torch.empty((8 * k, c, self.n, self.n), dtype=torch.float32, device=dev)
            pi = torch.empty((8 * k, self.policy_size), dtype=torch.float32, device=dev)
            z = torch.empty(8 * k, dtype=torch.float32, device=dev)
            torch.cuda.synchronize(dev)
            fp = C.POINTER(C.c_float)
            check(self.lib.examples_to_tensors(self._h, arr, k, C.cast(inputs.data_ptr(), fp), C.cast(pi.data_ptr(), fp),
                                               C.cast(z.data_ptr(), fp), 1))
            self.sync()
            return inputs, pi, z
        inputs = np.zeros((8 * k, c, self.n, self.n), dtype=np.float32)
        pi = np.zeros((8 * k, self.policy_size), dtype=np.float32)
        z = np.zeros(8 * k, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        check(self.lib.examples_to_tensors(self._h, arr, k, inputs.ctypes.data_as(fp), pi.ctypes.data_as(fp),
                                           z.ctypes.data_as(fp), 0))
        return inputs, pi, z


# ---- alpha_tak::Example text format / tak::Symmetry (host side, cold) ------------------------------------------------
def example_format(record: ReplayRecord) -> str:
    """`Display for Example<N>` (alpha-tak/src/example.rs:81-100)."""
    buf = C.create_string_buffer(1 << 14)
    check(_lib.load().tak_example_format(C.byref(record), buf, len(buf)))
    return buf.value.decode()


def example_parse(text: str, n: int) -> ReplayRecord:
    """`FromStr for Example<N>` (example.rs:102-133)."""
    rec = ReplayRecord()
    check(_lib.load().tak_example_parse(n, text.encode(), C.byref(rec)))
    return rec


def symmetry_move(move: int, n: int, k: int) -> int:
    out = C.c_uint16()
    check(_lib.load().tak_symmetry_move(n, move, k, C.byref(out)))
    return out.value


def symmetry_state(state: TakState, k: int) -> TakState:
    out = TakState()
    check(_lib.load().tak_symmetry_state(C.byref(state), k, C.byref(out)))
    return out


# ---- single-object mirrors of the Rust types --------------------------------------------------------------------
class _Slots:
    """Lazily created per-board-size engines whose game slots back `Game` objects."""

    engines = {}
    free = {}
    SLOTS = 256

    @classmethod
    def acquire(cls, n: int):
        if n not in cls.engines:
            cls.engines[n] = Engine(n, cls.SLOTS)
            cls.free[n] = list(range(cls.SLOTS - 1, -1, -1))
        if not cls.free[n]:
            raise RuntimeError("out of Game slots")
        return cls.engines[n], cls.free[n].pop()

    @classmethod
    def release(cls, n: int, slot: int):
        if n in cls.free:
            cls.free[n].append(slot)


class Game:
    """Mirror of `tak::Game<N>` (reference: tak/src/game.rs) whose state lives on the GPU.

    >>> g = Game.from_ptn_moves(5, ["d3", "c3", "c4", "1d3<", "1c4-", "Sc4"])
    >>> len(g.possible_moves())   # tak/tests/perft.rs:21-24
    87
    """

    def __init__(self, n: int = 5, half_komi: int = 0):
        self.n = n
        self.engine, self.slot = _Slots.acquire(n)
        self.engine.reset(self.slot, 1, half_komi)

    def __del__(self):
        try:
            _Slots.release(self.n, self.slot)
        except Exception:
            pass

    @classmethod
    def default(cls, n: int) -> "Game":
        return cls(n, 0)

    @classmethod
    def with_komi(cls, n: int, komi: int) -> "Game":
        return cls(n, komi * 2)

    @classmethod
    def with_half_komi(cls, n: int, half_komi: int) -> "Game":
        return cls(n, half_komi)

    @classmethod
    def from_ptn_moves(cls, n: int, moves: Iterable[str], half_komi: int = 0) -> "Game":
        g = cls(n, half_komi)
        for m in moves:
            st = g.play(m)
            if st != 0:
                raise TakNativeError(st, f"PlayError while playing {m}")
        return g

    @classmethod
    def from_state(cls, state: TakState) -> "Game":
        g = cls(state.n, state.half_komi)
        g.engine.upload([g.slot], [state])
        return g

    def clone(self) -> "Game":
        return Game.from_state(self.state())

    def play(self, move) -> int:
        if isinstance(move, str):
            move = parse_move(move, self.n)
        return int(self.engine.play([self.slot], [move])[0])

    def possible_moves(self) -> List[int]:
        return [int(m) for m in self.engine.possible_moves([self.slot])[0]]

    def result(self) -> int:
        return int(self.engine.result([self.slot])[0])

    def state(self) -> TakState:
        return self.engine.download([self.slot])[0]

    def set_half_komi(self, half_komi: int):
        s = self.state()
        s.half_komi = half_komi
        self.engine.upload([self.slot], [s])

    def tps(self) -> str:
        return tps_format(self.state())

    def perft(self, depth: int) -> int:
        return self.engine.perft(self.state(), depth)

    def safe_play(self, move) -> "Game":
        """`Game::safe_play` (game.rs:111-117): plays the move and returns the PRE-move backup; on a PlayError the game is
        restored from the backup and the error raised."""
        backup = self.clone()
        st = self.play(move)
        if st != 0:
            self.engine.upload([self.slot], [backup.state()])
            raise TakNativeError(st, "PlayError")
        return backup

    # ---- tak::Board accessors (board.rs:61-75), computed on the host from the downloaded state ----
    def board_full(self) -> bool:
        s = self.state()
        return all(s.height[i] > 0 for i in range(self.n * self.n))

    def flat_diff(self) -> int:
        """White flats on top minus black flats on top (board.rs:65-75)."""
        s, d = self.state(), 0
        for i in range(self.n * self.n):
            h = s.height[i]
            if h and s.top[i] == 0:
                hi = h - 1
                black = ((s.stack_lo[i] >> hi) & 1) if hi < 64 else ((s.stack_hi[i] >> (hi - 64)) & 1)
                d += -1 if black else 1
        return d

    def symmetries(self) -> List["Game"]:
        """`Symmetry::symmetries` for a game (symm.rs:82-97): the 8 transformed games, in the reference's order."""
        st = self.state()
        return [Game.from_state(symmetry_state(st, k)) for k in range(8)]


def default_starting_stones(n: int) -> Tuple[int, int]:
    """(stones, capstones) per player (tak/src/game.rs:10-20)."""
    return {3: (10, 0), 4: (15, 0), 5: (21, 1), 6: (30, 1), 7: (40, 2), 8: (50, 2)}[n]


class Player:
    """Mirror of `alpha_tak::Player` (alpha-tak/src/player.rs:23-199) for ONE game slot of an engine.

    The reference pipelines: a helper thread selects the NEXT batch of leaves (`request_batch`) while the caller
    evaluates and backs up the PREVIOUS one (`consume_batch`), so a backup always lands after the following batch was
    selected.  That order is kept deterministically here (the thread race itself is not reproduced): the
    engine queues the new batch behind the outstanding one and `mcts_devirtualize_first` backs up only the older.
    The game lives in slot `gid` of `engine`; `play_move` advances both the tree and the slot.
    """

    def __init__(self, engine: Engine, gid: int, batch: int, save_examples: bool = False,
                 state: Optional[TakState] = None, create_analysis: bool = False):
        from .analysis import Analysis
        self.engine, self.gid, self.batch, self.save_examples = engine, gid, batch, save_examples
        self.create_analysis = create_analysis
        self.examples: List[Tuple[TakState, List[Tuple[int, int]]]] = []
        engine.reserve_pending(2 * batch)
        if state is not None:
            engine.upload([gid], [state])
        engine.tree_reset([gid])
        st = engine.download([gid])[0]
        self.analysis = Analysis(engine.n, int(st.half_komi), int(st.ply))   # player.rs:58
        self._request_batch()                      # player.rs:66-67

    # Exactly one batch is outstanding between calls (every method below ends with a request), so "consume" is: back up
    # every leaf this game has queued.
    def _request_batch(self):                      # player.rs:98-100 (+ the rollout thread, :71-96)
        self.engine.virtual_rollout([self.gid], self.batch)   # terminal leaves need no evaluation (:83-87)

    def _consume_batch(self):                      # player.rs:102-110
        self.engine.devirtualize_first([self.gid], [1 << 30])

    def rollout(self, reps: int = 1):              # player.rs:130-133, `reps` calls fused into one ABI call
        self.engine.player_rollouts([self.gid], self.batch, reps)

    def add_noise(self, alpha: float, ratio: float, seed: int = 0):   # player.rs:123-127
        self._consume_batch()
        self.engine.apply_dirichlet([self.gid], alpha, ratio, seed)
        self._request_batch()

    def debug(self, depth: int = 10):                         # player.rs:113-115
        """NodeDebugInfo of the root: children sorted by visits, each with its principal continuation."""
        return self.engine.debug(self.gid, depth)

    def pick_move(self, exploitation: bool = True, rng: Optional[np.random.Generator] = None) -> int:   # player.rs:136-138
        if exploitation:
            return int(self.engine.pick_move([self.gid])[0])
        mv, vis, _, _ = self.engine.children(self.gid)        # play.rs:60-65: visit-weighted sample (thread_rng there)
        rng = rng or np.random.default_rng()
        return int(rng.choice(mv, p=vis / vis.sum()))

    def play_move(self, move: int, with_info: bool = True):   # player.rs:141-171
        self._consume_batch()                      # "rollout stale paths"
        if self.save_examples and with_info:
            mv, vis, _, _ = self.engine.children(self.gid)
            self.examples.append((self.engine.download([self.gid])[0], list(zip(mv.tolist(), vis.tolist()))))
        if self.create_analysis:                   # player.rs:155-161
            from .analysis import MAX_BRANCH_LENGTH
            if with_info:
                self.analysis.update(self.engine.debug(self.gid, MAX_BRANCH_LENGTH), move)
            else:
                self.analysis.add_move_without_info(move)
        self.engine.tree_play([self.gid], [move])
        if self.engine.play([self.gid], [move]).any():
            raise TakNativeError(-1, "play_move: illegal move")
        self._request_batch()

    def get_analysis(self):                                    # player.rs:196-198
        from .analysis import Analysis
        out, self.analysis = self.analysis, Analysis(self.engine.n, 0, 0)
        return out

    def get_examples(self, result: int) -> List[ReplayRecord]:    # player.rs:175-194
        if (result & 3) == RESULT_ONGOING:
            raise ValueError("cannot complete examples with ongoing game")
        white = 1.0 if (result & 3) == RESULT_WHITE else -1.0 if (result & 3) == RESULT_BLACK else 0.0
        out = []
        for state, policy in self.examples:
            r = ReplayRecord()
            r.state, r.n_children = state, len(policy)
            r.result = white if state.to_move == 0 else -white
            for i, (m, v) in enumerate(policy):
                r.moves[i], r.visits[i] = m, v
            out.append(r)
        self.examples = []
        return out
