"""Host-side mirror of `alpha_tak::{MoveInfo, NodeDebugInfo, Analysis}` -- the text the `analysis` and `playtak`
front-ends print from a search tree (alpha-tak/src/search/debug.rs:42-106, alpha-tak/src/analysis.rs:11-262).

The numbers come from the device tree through `mcts_debug` (one MoveInfo per root child incl. its principal
continuation); everything here is cold string formatting, kept byte-compatible with the reference's Display impls
(the reference pins one Analysis string in analysis.rs:265-288; tests/test_analysis_cpu.py holds it).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import MoveInfoRecord

MAX_BRANCH_LENGTH = 10          # analysis.rs:7
BRANCH_MIN_VISITS = 10_000      # analysis.rs:8
CANDIDATE_MOVE_RATIO = 0.9      # analysis.rs:9

_f32 = np.float32


def _move_str(move: int, n: int) -> str:
    from .engine import format_move
    return format_move(move, n)


@dataclass
class MoveInfo:
    """debug.rs:71-78.  `continuation` = [(move, visits of the node it leads to)]."""
    mov: int
    visits: int
    reward: float
    policy: float
    continuation: List[Tuple[int, int]] = field(default_factory=list)
    n: int = 6

    @classmethod
    def from_record(cls, r: MoveInfoRecord, n: int) -> "MoveInfo":
        k = int(r.cont_len)
        return cls(int(r.move), int(r.visits), float(r.reward), float(r.policy),
                   [(int(r.cont_moves[i]), int(r.cont_visits[i])) for i in range(k)], n)

    def ptn_comment(self, flip_reward: bool) -> str:          # debug.rs:81-84
        ev = -self.reward if flip_reward else self.reward
        return f" {{r: {ev:+.3f}, p: {self.policy:.4f}, v: {self.visits}}}"

    def __str__(self) -> str:                                 # debug.rs:87-105
        cont = " ".join(_move_str(m, self.n) for m, _ in self.continuation)
        return f"{_move_str(self.mov, self.n): <8} {self.visits: >8} {self.reward: >+8.4f} {self.policy: >8.4f} | {cont}\n"


class NodeDebugInfo:
    """debug.rs:42-69: root children in descending order of visits."""

    def __init__(self, moves: Sequence[MoveInfo]):
        self.moves = list(moves)

    def eval(self) -> float:                                  # debug.rs:47-54, f32 arithmetic in the same order
        total = _f32(sum(m.visits for m in self.moves) & 0xFFFFFFFF)
        acc = _f32(0.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            for m in self.moves:
                acc = _f32(acc + _f32(_f32(m.reward) * _f32(_f32(m.visits) / total)))
        return float(acc)

    def maybe_flip(self, flip: bool) -> "NodeDebugInfo":     # debug.rs:56-61
        if flip:
            for m in self.moves:
                m.reward = -m.reward
        return self

    def format(self, precision: Optional[int] = None) -> str:   # debug.rs:64-77 ("{:.k}" prints the top k moves)
        if not self.moves:
            return "Node has no children"
        out = f"evaluation: {self.eval():+.4f}\n"
        out += "turn      visited   reward   policy | continuation\n"
        for m in self.moves[:precision]:
            out += str(m)
        return out

    __str__ = format


_MARKS = {"blunder": "??", "mistake": "?", "strong": "!", "brilliancy": "!!"}   # analysis.rs:236-253


class Analysis:
    """analysis.rs:11-233: the annotated PTN a Player writes while a game is played."""

    def __init__(self, board_size: int, half_komi: int, start_ply: int):      # analysis.rs:23-35
        # Rust integer division / remainder truncate toward zero
        q = int(half_komi / 2)
        komi = str(q) + ("" if half_komi - 2 * q == 0 else ".5")
        self.settings = f'[Size "{board_size}"]\n[Komi "{komi}"]\n'
        self.n = board_size
        self.start_ply = start_ply
        self.played_moves: List[int] = []
        self.move_info: List[Optional[MoveInfo]] = []
        self.branches: List[Tuple[int, MoveInfo]] = []
        self.evals: List[float] = []
        self.marks: List[Tuple[int, str]] = []

    def add_setting(self, name: str, value) -> None:          # analysis.rs:37-39
        self.settings += f'[{name} "{value}"]\n'

    def add_move_without_info(self, mov: int) -> None:        # analysis.rs:41-44
        self.played_moves.append(mov)
        self.move_info.append(None)

    def add_move(self, mov: int, info: MoveInfo, ev: float) -> None:   # analysis.rs:46-50
        self.played_moves.append(mov)
        self.move_info.append(info)
        self.evals.append(ev)

    def update(self, debug_info: NodeDebugInfo, played_move: int) -> None:   # analysis.rs:52-92 (node.debug(10) passed in)
        ply = self.start_ply + len(self.played_moves)
        top_visits = debug_info.moves[0].visits if debug_info.moves else 0
        ev = debug_info.eval()
        if self.evals:
            diff = float(-(_f32(ev) + _f32(self.evals[-1])))      # due to flipping perspectives
            if diff <= -0.4:
                self.marks.append((ply - 1, "blunder"))
            elif -0.4 <= diff <= -0.15:
                self.marks.append((ply - 1, "mistake"))
            elif 0.1 <= diff <= 0.3:
                self.marks.append((ply - 1, "strong"))
            elif diff >= 0.3:
                self.marks.append((ply - 1, "brilliancy"))
        for info in debug_info.moves:
            if info.mov == played_move:
                self.add_move(played_move, info, ev)
                continue
            if float(_f32(info.visits)) > float(_f32(_f32(top_visits) * _f32(CANDIDATE_MOVE_RATIO))):
                self.branches.append((ply, info))

    def without_branches(self) -> "Analysis":                 # analysis.rs:94-97
        self.branches = []
        return self

    def __str__(self) -> str:                                 # analysis.rs:100-197
        out = self.settings
        moves = iter(self.played_moves)
        infos = iter(self.move_info)
        evals = iter(self.evals)
        marks = list(self.marks)
        mi = 0
        ply = self.start_ply
        next(evals, None)            # consume the first eval, so that a move is annotated with the eval AFTER it

        def mv(m):
            return _move_str(m, self.n)

        def mark_at(p):
            nonlocal mi
            if mi < len(marks) and marks[mi][0] == p:
                mi += 1
                return _MARKS[marks[mi - 1][1]]
            return ""

        def annotate(flip_eval: bool, flip_comment: bool) -> str:
            s = ""
            info = next(infos, None)
            if info is not None:
                e = next(evals, None)
                if e is not None:
                    s += f"{{evaluation: {e * (-1.0 if flip_eval else 1.0):+.3f}}}"
                s += info.ptn_comment(flip_comment)
            return s

        if self.start_ply % 2 != 0:
            out += f"{ply // 2 + 1}. -- "
            black = next(moves, None)
            if black is not None:
                out += mv(black) + mark_at(ply) + annotate(False, True)
            out += "\n"
            ply += 1
        while True:
            white = next(moves, None)
            if white is None:
                break
            out += f"{ply // 2 + 1}. " + mv(white) + mark_at(ply) + annotate(True, False) + " "
            ply += 1
            black = next(moves, None)
            if black is not None:
                out += mv(black) + mark_at(ply) + annotate(False, True)
            out += "\n"
            ply += 1
        for bply, branch in self.branches:
            out += "\n" + self._format_branch(bply, branch)
        return out

    def _format_branch(self, ply: int, info: MoveInfo) -> str:   # analysis.rs:199-234
        out = f"{{{ply}_{_move_str(info.mov, self.n)}}}\n"
        rest = [_move_str(m, self.n) for m, v in info.continuation if v > BRANCH_MIN_VISITS]
        it = iter(rest)
        move_num = 1 + ply // 2
        if ply % 2 == 0:
            out += f"{move_num}. {_move_str(info.mov, self.n)} {info.ptn_comment(False)} {next(it, '')}\n"
        else:
            out += f"{move_num}. -- {_move_str(info.mov, self.n)} {info.ptn_comment(True)}\n"
        move_num += 1
        for white in it:
            out += f"{move_num}. {white} {next(it, '')}\n"
            move_num += 1
        return out
