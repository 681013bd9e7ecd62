"""tak_b200: B200-native AlphaTak self-play engine (hot path of ViliamVadocz/tak) behind a C ABI.

The package is a thin host-side mirror of the reference interface over libtaknative.so (CUDA, sm_100a).
Importing it never touches oracle/ and never falls back to a CPU path: without the built library or a
CUDA device every compute call raises.
"""
from ._lib import LIB_PATH, ReplayRecord, SelfplayStats, TakNativeError, TakState, load  # noqa: F401
from .engine import (  # noqa: F401
    RESULT_BLACK, RESULT_DRAW, RESULT_FLAG, RESULT_ONGOING, RESULT_WHITE, Engine, Game, Player, default_starting_stones,
    format_move,
    boards_per_tile, example_format, example_parse, input_channels, move_index, parse_move, policy_size, state_init, symmetry_move,
    symmetry_state, tps_format, tps_parse,
)

from .analysis import Analysis, MoveInfo, NodeDebugInfo  # noqa: E402,F401
from .pit import PitResult, PlayerBatch, pit  # noqa: E402,F401

__all__ = ["default_starting_stones", "Analysis", "MoveInfo", "NodeDebugInfo", "PitResult", "PlayerBatch", "pit", "Engine", "Game", "Player", "TakState", "TakNativeError", "parse_move", "format_move", "move_index",
           "policy_size", "input_channels", "state_init", "tps_format", "tps_parse", "load", "LIB_PATH",
           "example_format", "example_parse", "symmetry_move", "symmetry_state"]
