"""ctypes binding of libtaknative.so (include/taknative.h).  Loading fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtaknative.so")

TAK_REPLAY_MAX_CHILDREN = 512


class TakState(C.Structure):
    """`tak_state_t`: POD mirror of tak::Game<N> (reference: tak/src/game.rs:25-35)."""

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

    def key(self) -> bytes:
        return bytes(self)


class EngineConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("n", C.c_int32),
        ("max_games", C.c_int32),
        ("nodes_per_game", C.c_int32),
        ("max_batch", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class SelfplayConfig(C.Structure):
    _fields_ = [
        ("rollouts", C.c_int32),
        ("half_komi", C.c_int32),
        ("instant_win", C.c_int32),
        ("exploit_ply", C.c_int32),
        ("noise_ply", C.c_int32),
        ("noise_alpha", C.c_float),
        ("noise_ratio", C.c_float),
        ("seed", C.c_uint64),
        ("max_plies", C.c_int32),
        ("game_id_base", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class SelfplayStats(C.Structure):
    _fields_ = [
        ("plies_played", C.c_uint64),
        ("games_completed", C.c_uint64),
        ("rollouts", C.c_uint64),
        ("evals", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("records", C.c_uint64),
        ("device_ms", C.c_double),
        ("net_ms", C.c_double),
        ("records_truncated", C.c_uint64),
        ("reserved", C.c_uint64 * 3),
    ]


TAK_DEBUG_MAX_DEPTH = 16


class MoveInfoRecord(C.Structure):
    """tak_move_info_t: one root child of Node::debug (search/debug.rs:9-24)."""
    _fields_ = [
        ("move", C.c_uint16),
        ("cont_len", C.c_uint16),
        ("visits", C.c_uint32),
        ("reward", C.c_float),
        ("policy", C.c_float),
        ("cont_moves", C.c_uint16 * TAK_DEBUG_MAX_DEPTH),
        ("cont_visits", C.c_uint32 * TAK_DEBUG_MAX_DEPTH),
    ]


class ReplayRecord(C.Structure):
    _fields_ = [
        ("game_id", C.c_int32),
        ("game_serial", C.c_int32),
        ("result", C.c_float),
        ("n_children", C.c_int32),
        ("state", TakState),
        ("moves", C.c_uint16 * TAK_REPLAY_MAX_CHILDREN),
        ("visits", C.c_uint32 * TAK_REPLAY_MAX_CHILDREN),
    ]


# every symbol include/taknative.h declares: name -> (restype, argtypes)
_vp, _i32, _u16, _u64, _f32 = C.c_void_p, C.c_int32, C.c_uint16, C.c_uint64, C.c_float
_P = C.POINTER
SYMBOLS = {
    "tak_engine_create": (_i32, [_P(EngineConfig), _P(_vp)]),
    "tak_engine_destroy": (_i32, [_vp]),
    "tak_engine_sync": (_i32, [_vp]),
    "tak_last_error": (C.c_char_p, []),
    "tak_version": (_i32, []),
    "tak_games_reset": (_i32, [_vp, _i32, _i32, _i32]),
    "tak_games_upload": (_i32, [_vp, _P(_i32), _i32, _P(TakState)]),
    "tak_games_download": (_i32, [_vp, _P(_i32), _i32, _P(TakState)]),
    "tak_possible_moves": (_i32, [_vp, _P(_i32), _i32, _P(_u16), _P(_i32), _i32]),
    "tak_play": (_i32, [_vp, _P(_i32), _P(_u16), _i32, _P(_i32)]),
    "tak_result": (_i32, [_vp, _P(_i32), _i32, _P(C.c_uint8)]),
    "tak_perft": (_i32, [_vp, _P(TakState), _i32, _P(_u64)]),
    "tak_perft_multi": (_i32, [_vp, _P(TakState), _i32, _i32, _P(_u64)]),
    "tak_perft_stats": (_i32, [_vp, _P(C.c_double), _P(_u64), _P(_u64)]),
    "tak_perft_profile": (_i32, [_vp, _P(C.c_double)]),
    "tak_playouts": (_i32, [_vp, _i32, _i32, _u64, _i32, _i32, _i32, _P(_i32), _P(C.c_uint8), _P(_u64),
                            _P(C.c_double)]),
    "tak_move_index": (_i32, [_i32, _u16, _P(_i32)]),
    "tak_policy_size": (_i32, [_i32, _P(_i32)]),
    "tak_ptn_parse": (_i32, [_i32, C.c_char_p, _P(_u16)]),
    "tak_ptn_format": (_i32, [_i32, _u16, C.c_char_p, _i32]),
    "tak_tps_format": (_i32, [_P(TakState), C.c_char_p, _i32]),
    "tak_tps_parse": (_i32, [_i32, C.c_char_p, _P(TakState)]),
    "tak_state_init": (_i32, [_i32, _i32, _P(TakState)]),
    "net_create": (_i32, [_vp, _i32]),
    "net_weights_size": (_i32, [_vp, _P(C.c_int64)]),
    "net_load_weights": (_i32, [_vp, _P(_f32), C.c_int64]),
    "net_load_weights_device": (_i32, [_vp, _vp, C.c_int64]),
    "net_input_channels": (_i32, [_i32, _P(_i32)]),
    "net_boards_per_tile": (_i32, [_i32, _P(_i32)]),
    "net_game_repr": (_i32, [_vp, _P(TakState), _i32, _P(_f32)]),
    "net_policy_eval": (_i32, [_vp, _P(TakState), _i32, _P(_f32), _P(_f32)]),
    "net_policy_logits": (_i32, [_vp, _P(TakState), _i32, _P(_f32), _P(_f32)]),
    "tak_host_alloc": (_i32, [C.c_size_t, _P(_vp)]),
    "tak_host_free": (_i32, [_vp]),
    "net_forward_timed": (_i32, [_vp, _i32, _i32, _i32, _P(C.c_double)]),
    "net_forward_profile": (_i32, [_vp, _i32, _i32, _i32, _P(C.c_double)]),
    "net_train_begin": (_i32, [_vp, _i32]),
    "net_train_chunk": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _P(_f32)]),
    "net_train_step": (_i32, [_vp, _f32, _f32]),
    "net_train_get": (_i32, [_vp, _i32, _P(_f32), C.c_int64]),
    "net_train_grad_ptr": (_i32, [_vp, _P(_vp), _P(C.c_int64)]),
    "net_train_stats": (_i32, [_vp, _P(C.c_double), _P(_i32), _P(_i32)]),
    "net_train_end": (_i32, [_vp]),
    "mcts_tree_reset": (_i32, [_vp, _P(_i32), _i32]),
    "mcts_virtual_rollout": (_i32, [_vp, _P(_i32), _i32, _i32]),
    "mcts_pending": (_i32, [_vp, _P(_i32), _P(_i32), _P(TakState), _i32]),
    "mcts_devirtualize": (_i32, [_vp]),
    "mcts_devirtualize_with": (_i32, [_vp, _P(_f32), _P(_f32), _i32]),
    "mcts_rollouts": (_i32, [_vp, _P(_i32), _i32, _i32]),
    "mcts_player_rollouts": (_i32, [_vp, _P(_i32), _i32, _i32, _i32]),
    "mcts_children": (_i32, [_vp, _i32, _P(_u16), _P(C.c_uint32), _P(_f32), _P(_f32), _i32, _P(_i32)]),
    "mcts_children_batch": (_i32, [_vp, _P(_i32), _i32, _P(_u16), _P(C.c_uint32), _P(_i32), _i32]),
    "mcts_root": (_i32, [_vp, _i32, _P(C.c_uint32), _P(C.c_uint32), _P(_f32)]),
    "mcts_debug": (_i32, [_vp, _i32, _i32, _P(MoveInfoRecord), _i32, _P(_i32)]),
    "mcts_pick_move": (_i32, [_vp, _P(_i32), _i32, _P(_u16)]),
    "mcts_pick_move_sampled": (_i32, [_vp, _P(_i32), _i32, _u64, _P(_u16)]),
    "mcts_play": (_i32, [_vp, _P(_i32), _P(_u16), _i32]),
    "mcts_apply_dirichlet": (_i32, [_vp, _P(_i32), _i32, _f32, _f32, _u64]),
    "selfplay_begin": (_i32, [_vp, _P(SelfplayConfig)]),
    "selfplay_step": (_i32, [_vp, _i32, _P(SelfplayStats)]),
    "selfplay_drain": (_i32, [_vp, _P(ReplayRecord), _i32, _P(_i32)]),
    "mcts_reserve_pending": (_i32, [_vp, _i32]),
    "mcts_devirtualize_first": (_i32, [_vp, _P(_i32), _i32, _P(_i32)]),
    "tak_example_format": (_i32, [_P(ReplayRecord), C.c_char_p, _i32]),
    "tak_example_parse": (_i32, [_i32, C.c_char_p, _P(ReplayRecord)]),
    "tak_symmetry_move": (_i32, [_i32, _u16, _i32, _P(_u16)]),
    "tak_symmetry_state": (_i32, [_P(TakState), _i32, _P(TakState)]),
    "examples_to_tensors": (_i32, [_vp, _P(ReplayRecord), _i32, _P(_f32), _P(_f32), _P(_f32), _i32]),
    "tak_comm_unique_id": (_i32, [_P(C.c_uint8), _i32]),
    "tak_comm_init": (_i32, [_vp, _P(C.c_uint8), _i32, _i32]),
    "tak_comm_destroy": (_i32, [_vp]),
    "tak_comm_info": (_i32, [_vp, _P(_i32), _P(_i32), _P(_u64)]),
    "net_broadcast_weights": (_i32, [_vp, _P(_f32), C.c_int64, _i32]),
    "selfplay_gather_replay": (_i32, [_vp, _P(ReplayRecord), _i32, _P(ReplayRecord), _i32, _P(_i32)]),
    "net_train_allreduce": (_i32, [_vp]),
    "tak_comm_sum_u64": (_i32, [_vp, _P(_u64), _i32]),
    "tak_comm_max_f64": (_i32, [_vp, _P(C.c_double), _i32]),
}

_lib = None


class TakNativeError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"taknative error {code}: {text}")
        self.code = code


def _preload_nccl() -> None:
    """libtaknative.so needs libnccl.so.2.  PyTorch bundles its own (newer) NCCL under site-packages/nvidia/nccl and loads
    it by that SONAME too; whichever is loaded first serves both.  Load the bundled one first when it exists so that a
    later `import torch` in the same process finds the version it was built against; otherwise the system library that
    the linker recorded is used."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        paths = list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []
        for base in paths:
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def load() -> C.CDLL:
    """dlopen the in-tree shared library and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C tak_b200/csrc).  tak_b200 has no CPU fallback."
            )
        _preload_nccl()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code: int) -> int:
    if code < 0:
        raise TakNativeError(code, load().tak_last_error().decode(errors="replace"))
    return code
