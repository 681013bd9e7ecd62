"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

ctypes front-end of the CPU oracle (oracle/liboracle.so, built from tak_oracle.hpp / alphatak_oracle.hpp,
a literal C++ restatement of the reference's `tak` and `alpha-tak::search/repr` crates).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


class TakState(C.Structure):
    """POD mirror of tak::Game<N>; identical to `tak_state_t` in include/taknative.h."""

    _fields_ = [
        ("n", C.c_uint8),
        ("to_move", C.c_uint8),
        ("ply", C.c_uint16),
        ("white_stones", C.c_uint8),
        ("white_caps", C.c_uint8),
        ("black_stones", C.c_uint8),
        ("black_caps", C.c_uint8),
        ("half_komi", C.c_int8),
        ("reversible_plies", C.c_uint8),
        ("_pad", C.c_uint8 * 6),
        ("height", C.c_uint8 * 64),
        ("top", C.c_uint8 * 64),
        ("stack_lo", C.c_uint64 * 64),
        ("stack_hi", C.c_uint64 * 64),
    ]

    def key(self):
        return bytes(self)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "tak_oracle.hpp", "alphatak_oracle.hpp")]
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None
_native = False


def _cpu_tag() -> str:
    """Identifies this machine's CPU (model + ISA flags): a -march=native library must never run on another one."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            lines = [ln for ln in f if ln.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("".join(lines).encode()).hexdigest()[:10]
    except OSError:
        return "unknown"


_native_path = None


def use_native_build() -> bool:
    """bench.py's CPU-baseline legs: compile the oracle with -march=native on THIS machine (file name keyed by the CPU, so
    a library built elsewhere and carried along in the tree is never loaded) and use it.  Must be called before the
    first oracle call; falls back to the portable build if the compiler is missing."""
    global _native, _native_path
    if _lib is not None:
        return _native
    path = os.path.join(_HERE, f"liboracle_native_{_cpu_tag()}.so")
    try:
        if not os.path.exists(path):
            subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread",
                                   "-shared", "-o", path, os.path.join(_HERE, "oracle_capi.cpp")])
        _native, _native_path = True, path
    except Exception:
        _native = False
    return _native


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_native_path if _native else _LIB_PATH)
        vp, i32, u16, u64, f32 = C.c_void_p, C.c_int, C.c_uint16, C.c_uint64, C.c_float
        sig = {
            "orc_game_new": (vp, [i32, i32]),
            "orc_game_free": (None, [vp]),
            "orc_game_clone": (vp, [vp]),
            "orc_game_play": (i32, [vp, u16]),
            "orc_game_moves": (i32, [vp, C.POINTER(u16), i32]),
            "orc_game_result": (i32, [vp]),
            "orc_game_get": (None, [vp, C.POINTER(TakState)]),
            "orc_game_from_state": (vp, [C.POINTER(TakState)]),
            "orc_game_set_half_komi": (None, [vp, i32]),
            "orc_game_flat_diff": (i32, [vp]),
            "orc_selfplay_instant_win": (i32, [vp, C.POINTER(u16), C.POINTER(C.c_uint32), i32, C.POINTER(i32)]),
            "orc_perft": (u64, [vp, i32]),
            "orc_perft_mt": (u64, [vp, i32, i32]),
            "orc_parse_move": (i32, [C.c_char_p, i32]),
            "orc_format_move": (i32, [u16, i32, C.c_char_p, i32]),
            "orc_game_from_ptn": (vp, [i32, i32, C.c_char_p, C.POINTER(i32)]),
            "orc_game_tps": (i32, [vp, C.c_char_p, i32]),
            "orc_game_from_tps": (vp, [C.c_char_p, i32]),
            "orc_input_channels": (i32, [i32]),
            "orc_board_channels": (i32, [i32]),
            "orc_policy_size": (i32, [i32]),
            "orc_output_size": (i32, [i32]),
            "orc_move_index": (i32, [u16, i32]),
            "orc_legacy_move_5": (i32, [i32, C.c_char_p, i32]),
            "orc_game_repr": (None, [vp, C.POINTER(f32)]),
            "orc_board_repr": (None, [vp, i32, C.POINTER(f32)]),
            "orc_search_new": (vp, []),
            "orc_search_free": (None, [vp]),
            "orc_search_virtual_rollout": (i32, [vp, vp]),
            "orc_search_pending": (i32, [vp]),
            "orc_search_pending_state": (i32, [vp, i32, C.POINTER(TakState)]),
            "orc_search_pending_path": (i32, [vp, i32, C.POINTER(C.c_int32), i32]),
            "orc_search_devirtualize": (i32, [vp, C.POINTER(f32), f32]),
            "orc_search_rollouts_dummy": (None, [vp, vp, i32]),
            "orc_search_children": (
                i32,
                [vp, i32, C.POINTER(u16), C.POINTER(C.c_uint32), C.POINTER(f32), C.POINTER(f32),
                 C.POINTER(C.c_uint32), i32],
            ),
            "orc_search_root": (None, [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(f32)]),
            "orc_search_debug": (i32, [vp, i32, i32, C.POINTER(u16), C.POINTER(C.c_uint32), C.POINTER(f32),
                                       C.POINTER(f32), C.POINTER(C.c_int32), C.POINTER(u16), C.POINTER(C.c_uint32),
                                       i32]),
            "orc_search_pick": (i32, [vp, i32]),
            "orc_search_play": (i32, [vp, u16, i32]),
            "orc_search_reset": (None, [vp]),
            "orc_search_apply_noise": (None, [vp, C.POINTER(f32), f32]),
            "orc_search_node_count": (u64, [vp]),
            "orc_symmetry_move": (i32, [u16, i32, i32]),
            "orc_symmetry_game": (None, [vp, i32, C.POINTER(TakState)]),
            "orc_example_to_tensors": (None, [vp, C.POINTER(u16), C.POINTER(C.c_uint32), i32, f32, C.POINTER(f32),
                                              C.POINTER(f32)]),
            "orc_example_format": (i32, [vp, C.POINTER(u16), C.POINTER(C.c_uint32), i32, f32, C.c_char_p, i32]),
            "orc_example_parse": (vp, [C.c_char_p, i32, C.POINTER(u16), C.POINTER(C.c_uint32), i32, C.POINTER(i32),
                                       C.POINTER(f32)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


RESULT_NAMES = {0: "Ongoing", 1: "WhiteFlat", 2: "BlackFlat", 3: "Draw", 0x11: "WhiteRoad", 0x12: "BlackRoad",
                0x13: "DrawReversible"}


def parse_move(text: str, n: int) -> int:
    m = lib().orc_parse_move(text.encode(), n)
    if m < 0:
        raise ValueError(f"bad PTN move {text!r}")
    return m


def format_move(move: int, n: int) -> str:
    buf = C.create_string_buffer(32)
    lib().orc_format_move(move, n, buf, 32)
    return buf.value.decode()


class Game:
    """Mirror of tak::Game<N> over the oracle (reference: tak/src/game.rs)."""

    def __init__(self, n: int = 5, half_komi: int = 0, _handle=None):
        self.n = n
        self._h = _handle if _handle is not None else lib().orc_game_new(n, half_komi)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_game_free(self._h)
            self._h = None

    @classmethod
    def with_komi(cls, n: int, komi: int) -> "Game":
        return cls(n, komi * 2)

    @classmethod
    def from_ptn_moves(cls, n: int, moves, half_komi: int = 0) -> "Game":
        st = C.c_int(0)
        text = " ".join(moves) if not isinstance(moves, str) else moves
        h = lib().orc_game_from_ptn(n, half_komi, text.encode(), C.byref(st))
        if not h:
            raise ValueError(f"from_ptn_moves failed with status {st.value}")
        return cls(n, _handle=h)

    @classmethod
    def from_tps(cls, n: int, tps: str) -> "Game":
        h = lib().orc_game_from_tps(tps.encode(), n)
        if not h:
            raise ValueError("bad TPS")
        return cls(n, _handle=h)

    @classmethod
    def from_state(cls, state: TakState) -> "Game":
        return cls(state.n, _handle=lib().orc_game_from_state(C.byref(state)))

    def clone(self) -> "Game":
        return Game(self.n, _handle=lib().orc_game_clone(self._h))

    def play(self, move) -> int:
        if isinstance(move, str):
            move = parse_move(move, self.n)
        return lib().orc_game_play(self._h, move)

    def possible_moves(self):
        buf = (C.c_uint16 * 4096)()
        k = lib().orc_game_moves(self._h, buf, 4096)
        return list(buf[:k])

    def flat_diff(self) -> int:
        return lib().orc_game_flat_diff(self._h)

    def result(self) -> int:
        return lib().orc_game_result(self._h)

    def state(self) -> TakState:
        s = TakState()
        lib().orc_game_get(self._h, C.byref(s))
        return s

    def set_half_komi(self, hk: int):
        lib().orc_game_set_half_komi(self._h, hk)

    def instant_win_policy(self):
        """self_play.rs:121-140: ([(move, 1000 if it wins on the spot else 1)], any win)."""
        mv, vis = (C.c_uint16 * 4096)(), (C.c_uint32 * 4096)()
        win = C.c_int(0)
        k = lib().orc_selfplay_instant_win(self._h, mv, vis, 4096, C.byref(win))
        return [(mv[i], vis[i]) for i in range(k)], bool(win.value)

    def perft(self, depth: int, threads: int = 1) -> int:
        if threads > 1:
            return lib().orc_perft_mt(self._h, depth, threads)
        return lib().orc_perft(self._h, depth)

    def tps(self) -> str:
        buf = C.create_string_buffer(4096)
        lib().orc_game_tps(self._h, buf, 4096)
        return buf.value.decode()

    def repr(self) -> np.ndarray:
        c = lib().orc_input_channels(self.n)
        out = np.zeros((c, self.n, self.n), dtype=np.float32)
        lib().orc_game_repr(self._h, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def board_repr(self, to_move: int) -> np.ndarray:
        c = lib().orc_board_channels(self.n)
        out = np.zeros((c, self.n, self.n), dtype=np.float32)
        lib().orc_board_repr(self._h, to_move, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out


class Search:
    """Mirror of alpha_tak::Node as a search root plus its queue of un-evaluated leaves."""

    def __init__(self, n: int):
        self.n = n
        self._h = lib().orc_search_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_search_free(self._h)
            self._h = None

    def virtual_rollout(self, game: Game) -> int:
        return lib().orc_search_virtual_rollout(self._h, game._h)

    def pending(self) -> int:
        return lib().orc_search_pending(self._h)

    def pending_state(self, idx: int = 0) -> TakState:
        s = TakState()
        if lib().orc_search_pending_state(self._h, idx, C.byref(s)) != 0:
            raise IndexError(idx)
        return s

    def pending_path(self, idx: int = 0):
        buf = (C.c_int32 * 1024)()
        k = lib().orc_search_pending_path(self._h, idx, buf, 1024)
        return list(buf[:k])

    def devirtualize(self, policy: np.ndarray, value: float):
        policy = np.ascontiguousarray(policy, dtype=np.float32)
        assert policy.size == lib().orc_policy_size(self.n)
        r = lib().orc_search_devirtualize(self._h, policy.ctypes.data_as(C.POINTER(C.c_float)), float(value))
        assert r == 0

    def rollouts_dummy(self, game: Game, count: int):
        lib().orc_search_rollouts_dummy(self._h, game._h, count)

    def children(self):
        cap = 4096
        mv = (C.c_uint16 * cap)()
        vis = (C.c_uint32 * cap)()
        pri = (C.c_float * cap)()
        rew = (C.c_float * cap)()
        virt = (C.c_uint32 * cap)()
        k = lib().orc_search_children(self._h, self.n, mv, vis, pri, rew, virt, cap)
        return (np.array(mv[:k], dtype=np.uint16), np.array(vis[:k], dtype=np.uint32),
                np.array(pri[:k], dtype=np.float32), np.array(rew[:k], dtype=np.float32),
                np.array(virt[:k], dtype=np.uint32))

    def debug(self, depth: int = 10):
        """Node::debug(depth): [(move, visits, reward, policy, [(move, visits), ...])] in descending order of visits."""
        cap = 4096
        mv, vis = (C.c_uint16 * cap)(), (C.c_uint32 * cap)()
        rew, pol = (C.c_float * cap)(), (C.c_float * cap)()
        clen = (C.c_int32 * cap)()
        cmv, cvis = (C.c_uint16 * (cap * 16))(), (C.c_uint32 * (cap * 16))()
        k = lib().orc_search_debug(self._h, self.n, depth, mv, vis, rew, pol, clen, cmv, cvis, cap)
        return [(mv[i], vis[i], rew[i], pol[i], [(cmv[i * 16 + j], cvis[i * 16 + j]) for j in range(min(16, clen[i]))])
                for i in range(k)]

    def root(self):
        v, vv, r = C.c_uint32(), C.c_uint32(), C.c_float()
        lib().orc_search_root(self._h, C.byref(v), C.byref(vv), C.byref(r))
        return v.value, vv.value, r.value

    def pick_move(self) -> int:
        return lib().orc_search_pick(self._h, self.n)

    def play(self, move: int):
        if lib().orc_search_play(self._h, move, self.n) != 0:
            raise ValueError("tried to play an invalid move")

    def reset(self):
        lib().orc_search_reset(self._h)

    def apply_noise(self, noise: np.ndarray, ratio: float):
        noise = np.ascontiguousarray(noise, dtype=np.float32)
        lib().orc_search_apply_noise(self._h, noise.ctypes.data_as(C.POINTER(C.c_float)), ratio)

    def node_count(self) -> int:
        return lib().orc_search_node_count(self._h)


def input_channels(n): return lib().orc_input_channels(n)
def policy_size(n): return lib().orc_policy_size(n)
def move_index(move, n): return lib().orc_move_index(move, n)


def legacy_moves_5():
    out = []
    buf = C.create_string_buffer(32)
    for i in range(1575):
        lib().orc_legacy_move_5(i, buf, 32)
        out.append(buf.value.decode())
    return out


# ---- tak::Symmetry / alpha_tak::Example (tak/src/symm.rs, alpha-tak/src/example.rs) ---------------------------------
def symmetry_move(move: int, n: int, k: int) -> int:
    """Symmetry::<N>::symmetries(move)[k]."""
    return lib().orc_symmetry_move(move, n, k)


def symmetry_game(game: "Game", k: int) -> TakState:
    """Symmetry::<N>::symmetries(game)[k] as a POD state."""
    s = TakState()
    lib().orc_symmetry_game(game._h, k, C.byref(s))
    return s


class Example:
    """Mirror of alpha_tak::Example<N>: (game, [(move, visits)], result)."""

    def __init__(self, game: "Game", policy, result: float):
        self.game, self.policy, self.result = game, [(int(m), int(v)) for m, v in policy], float(result)

    def _arrays(self):
        k = len(self.policy)
        mv = (C.c_uint16 * max(k, 1))(*[m for m, _ in self.policy])
        vis = (C.c_uint32 * max(k, 1))(*[v for _, v in self.policy])
        return mv, vis, k

    def to_tensors(self):
        """Example::to_tensors: (inputs [8, C, n, n], pi [8, policy_size], z [8])."""
        n = self.game.n
        c, p = lib().orc_input_channels(n), lib().orc_policy_size(n)
        inputs = np.zeros((8, c, n, n), dtype=np.float32)
        pi = np.zeros((8, p), dtype=np.float32)
        mv, vis, k = self._arrays()
        lib().orc_example_to_tensors(self.game._h, mv, vis, k, self.result,
                                     inputs.ctypes.data_as(C.POINTER(C.c_float)), pi.ctypes.data_as(C.POINTER(C.c_float)))
        return inputs, pi, np.full(8, self.result, dtype=np.float32)

    def __str__(self) -> str:
        buf = C.create_string_buffer(1 << 14)
        mv, vis, k = self._arrays()
        if lib().orc_example_format(self.game._h, mv, vis, k, self.result, buf, len(buf)) < 0:
            raise ValueError("example too long")
        return buf.value.decode()

    @classmethod
    def parse(cls, text: str, n: int) -> "Example":
        mv, vis = (C.c_uint16 * 512)(), (C.c_uint32 * 512)()
        cnt, res = C.c_int(0), C.c_float(0)
        h = lib().orc_example_parse(text.encode(), n, mv, vis, 512, C.byref(cnt), C.byref(res))
        if not h:
            raise ValueError("bad Example line")
        return cls(Game(n, _handle=h), [(mv[i], vis[i]) for i in range(cnt.value)], res.value)
