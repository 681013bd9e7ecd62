"""`train::pit` (train/src/pit.rs:15-96) on the device engines: the new network against the old one, every pit game
played CONCURRENTLY instead of one after the other.

The reference plays 128 openings x 2 colours sequentially, each game through two `Player`s (one per network) that both
follow the game; only the player to move searches (50 `Player::rollout`s of 16 pipelined virtual rollouts each).  Games
are independent, so here all of them live side by side in the slots of two engines (one per network) and every step is a
batched ABI call over the games concerned -- `PlayerBatch` is `alpha_tak::Player` (player.rs:98-171) for many slots at
once, with exactly the per-game schedule `tak_b200.Player` keeps (tests/test_pit_gpu.py compares them tree for tree).
The reference's early exit ("result is already known", pit.rs:19-22) depends on the sequential order and is not taken.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .engine import RESULT_BLACK, RESULT_ONGOING, RESULT_WHITE, Engine, parse_move

PIT_GAMES = 128        # pit.rs:5
BATCH_SIZE = 16        # pit.rs:6
ROLLOUTS = 50          # pit.rs:7
RANDOM_PLIES = 2       # pit.rs:9


class PlayerBatch:
    """`alpha_tak::Player` for a list of game slots of one engine, stepped in lock step."""

    def __init__(self, engine: Engine, gids: Sequence[int], batch: int):
        self.engine, self.batch = engine, batch
        self.gids = [int(g) for g in gids]
        engine.reserve_pending(2 * batch)
        engine.tree_reset(self.gids)
        self.request(self.gids)                           # player.rs:66-67

    # exactly one batch per game is outstanding between calls, so "consume" = back up everything the game has queued
    def request(self, gids: Sequence[int]):               # player.rs:98-100 + the rollout thread (:71-96)
        if len(gids):
            self.engine.virtual_rollout(gids, self.batch)

    def consume(self, gids: Sequence[int]):               # player.rs:102-110
        if len(gids):
            self.engine.devirtualize_first(gids, [1 << 30] * len(gids))

    def rollout(self, gids: Sequence[int], reps: int = 1):   # player.rs:130-133 (x reps, fused into one ABI call)
        if len(gids):
            self.engine.player_rollouts(gids, self.batch, reps)

    def pick_move(self, gids: Sequence[int]) -> np.ndarray:   # player.rs:136-138, exploitation = true
        return self.engine.pick_move(gids)

    def play_move(self, gids: Sequence[int], moves: Sequence[int]):   # player.rs:141-171
        if not len(gids):
            return
        self.consume(gids)                                # "rollout stale paths"
        self.engine.tree_play(gids, moves)
        if self.engine.play(gids, moves).any():
            raise ValueError("play_move: illegal move")
        self.request(gids)


@dataclass
class PitResult:                                          # pit.rs:98-129
    wins: int = 0
    losses: int = 0
    draws: int = 0

    def win_rate(self) -> float:
        return self.wins / (self.wins + self.losses) if self.wins + self.losses else float("nan")

    def update(self, result: int, color: int):
        r = result & 3
        if r in (RESULT_WHITE, RESULT_BLACK):
            if (0 if r == RESULT_WHITE else 1) == color:
                self.wins += 1
            else:
                self.losses += 1
        elif r != RESULT_ONGOING:
            self.draws += 1


def random_opening(engine: Engine, slot: int, rng: np.random.Generator, half_komi: int = 4) -> List[int]:
    """pit.rs:34-62: a1, then a<N>/<last file><N> at random, then RANDOM_PLIES uniformly random flat/cap placements."""
    n = engine.n
    engine.reset(slot, 1, half_komi)
    opening = [parse_move("a1", n), parse_move(f"a{n}" if rng.random() < 0.5 else f"{'abcdefgh'[n - 1]}{n}", n)]
    for m in opening:
        assert not engine.play([slot], [m]).any()
    for _ in range(RANDOM_PLIES):
        moves = engine.possible_moves([slot])[0]
        # placements have pattern byte 0; bits 6-7 carry the piece: flat 0 / wall 1 / cap 2 (include/taknative.h)
        cand = [int(m) for m in moves if (int(m) >> 8) == 0 and ((int(m) >> 6) & 3) != 1]
        m = cand[int(rng.integers(len(cand)))]
        opening.append(m)
        assert not engine.play([slot], [m]).any()
    return opening


def pit(new: Engine, old: Engine, games: int = PIT_GAMES, batch: int = BATCH_SIZE, rollouts: int = ROLLOUTS,
        seed: int = 0, half_komi: int = 4, max_plies: int = 400, log: Optional[list] = None) -> PitResult:
    """Play `games` openings x both colours, `new`'s network against `old`'s.  Both engines need 2*games slots and the
    same board size; their networks must be loaded.  Returns new's PitResult (pit.rs:15-96)."""
    assert new.n == old.n and new.max_games >= 2 * games and old.max_games >= 2 * games
    rng = np.random.default_rng(seed)
    total = 2 * games
    gids = list(range(total))
    colour_of_new = [g & 1 for g in gids]                 # game 2i: new plays White; game 2i+1: new plays Black
    openings = []
    for i in range(games):                                # slot 0 is scratch here: every slot is reset right after
        op = random_opening(new, 0, rng, half_komi)
        openings += [op, op]
    for eng in (new, old):
        eng.reset(0, total, half_komi)                    # Game::with_komi(2) (pit.rs:28)
    players = [PlayerBatch(new, gids, batch), PlayerBatch(old, gids, batch)]
    for ply in range(2 + RANDOM_PLIES):                   # pit.rs:64-68: with_info = false
        mv = [openings[g][ply] for g in gids]
        for p in players:
            p.play_move(gids, mv)
    result = PitResult()
    live = gids
    for _ in range(max_plies):
        res = new.result(live)
        for g, r in zip(live, res):
            if (int(r) & 3) != RESULT_ONGOING:
                result.update(int(r), colour_of_new[g])
                if log is not None:
                    log.append((g, int(r), int(new.download([g])[0].ply)))
        live = [g for g, r in zip(live, res) if (int(r) & 3) == RESULT_ONGOING]
        if not live:
            break
        to_move = [int(s.to_move) for s in new.download(live)]
        side = [[g for g, c in zip(live, to_move) if (c == colour_of_new[g]) == (k == 0)] for k in (0, 1)]
        picks = {}
        for k in (0, 1):                                  # the player to move searches (pit.rs:71-83)
            if side[k]:
                players[k].rollout(side[k], rollouts)
                for g, m in zip(side[k], players[k].pick_move(side[k])):
                    picks[g] = int(m)
        mv = [picks[g] for g in live]
        for p in players:                                 # pit.rs:84-86
            p.play_move(live, mv)
    return result
