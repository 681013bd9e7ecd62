"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU restatement of `train::self_play_parallel` (reference: train/src/self_play.rs:96-262) over the oracle's Game / Node
(oracle/tak_oracle.hpp, alphatak_oracle.hpp).  One call of `iteration()` is one pass of the reference's `while` body:

    forced opening of fresh slots (:108-117) -> "play winning moves if there are any" (:119-171: the scan itself is
    oracle.Game.instant_win_policy, C++) -> [Dirichlet noise (:173-180)] -> ROLLOUTS x (virtual rollout of every slot,
    one batched policy_eval, devirtualize) (:181-210) -> pick / IncompleteExample / Node::play / Game::play, finished
    games complete their examples and the slot restarts (:212-258).

What cannot be restated is injected by the caller: the two `thread_rng` draws (opening coin flip, sampled pick) and the
network.  Pure-Python loops: meant for a handful of slots and a few dozen rollouts (parity tests only).
Reference quirks kept: after an instant win the slot is reset to the empty board and is searched from ply 0 in the SAME
pass, so it never receives the forced opening (:148-152 vs :110-116).  The `completed_games + WORKERS <
SELF_PLAY_GAMES` tail (:149,237: slots retire near the end of a run) is a parameter: `total_games=None` never retires.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

import oracle


class Example:
    """alpha_tak::Example (example.rs:28-33) with the slot bookkeeping the device records carry."""

    def __init__(self, slot: int, serial: int, state, policy, result: float):
        self.slot, self.serial, self.state, self.policy, self.result = slot, serial, state, policy, result


def result_to_number(result: int) -> float:                       # self_play.rs:264-275
    r = result & 0xF
    if r == 1:
        return 1.0
    if r == 2:
        return -1.0
    if r == 3:
        return 0.0
    raise ValueError("cannot complete examples with ongoing game")


class SelfPlayParallel:
    def __init__(self, n: int, workers: int, rollouts: int,
                 policy_eval: Optional[Callable[[Sequence[oracle.TakState]], Tuple[np.ndarray, np.ndarray]]] = None,
                 coin: Optional[Callable[[int, int], bool]] = None, exploit_plies: int = 40, instant_win: bool = True,
                 komi: int = 2, total_games: Optional[int] = None,
                 sample: Optional[Callable[[int, int, list], int]] = None):
        self.n, self.workers, self.rollouts = n, workers, rollouts
        self.policy_eval, self.coin, self.sample = policy_eval, coin, sample
        self.exploit_plies, self.instant_win, self.komi, self.total_games = exploit_plies, instant_win, komi, total_games
        self.nodes = [oracle.Search(n) for _ in range(workers)]                       # :102
        self.games: List[Optional[oracle.Game]] = [oracle.Game.with_komi(n, komi) for _ in range(workers)]   # :103
        self.incomplete: List[list] = [[] for _ in range(workers)]                    # :104
        self.serial = [0] * workers          # games a slot has completed (device records carry it)
        self.completed_games = 0
        self.examples: List[Example] = []
        self.moves_played: List[Optional[int]] = [None] * workers   # the move each slot played in the last pass

    def _restart(self, i: int):
        self.nodes[i] = oracle.Search(self.n)                                         # *node = Node::default()
        if self.total_games is None or self.completed_games + self.workers < self.total_games:
            self.games[i] = oracle.Game.with_komi(self.n, self.komi)                  # :150 / :238
        else:
            self.games[i] = None
        self.serial[i] += 1

    def _complete(self, i: int, white_result: float):
        for state, policy in self.incomplete[i]:                                      # :157-164 / :246-253
            perspective = white_result if state.to_move == 0 else -white_result
            self.examples.append(Example(i, self.serial[i], state, policy, perspective))
        self.incomplete[i] = []

    def iteration(self):
        n = self.n
        # ---- play opening moves (:108-117)
        for i, game in enumerate(self.games):
            if game is None:
                continue
            if game.state().ply == 0:
                game.play("a1")
                first = self.coin(i, self.serial[i]) if self.coin else True
                game.play(f"a{n}" if first else f"{'abcdefgh'[n - 1]}{n}")
        # ---- play winning moves if there are any (:119-171)
        if self.instant_win:
            for i, game in enumerate(self.games):
                if game is None:
                    continue
                policy, win = game.instant_win_policy()
                if win:
                    self.incomplete[i].append((game.state(), policy))
                    self.completed_games += 1
                    white_result = 1.0 if game.state().to_move == 0 else -1.0       # Winner{color: to_move} (:146)
                    self._complete(i, white_result)
                    self._restart(i)                 # the fresh game stays at ply 0 for the rest of this pass
        # ---- noise at the start of a ply (:173-180): unreproducible RNG -> parity runs leave it off
        # ---- ROLLOUTS x (virtual rollouts, one batched evaluation, devirtualise) (:181-210)
        live = [i for i, g in enumerate(self.games) if g is not None]
        for _ in range(self.rollouts):
            pend = [i for i in live if self.nodes[i].virtual_rollout(self.games[i]) == 0]
            if not pend:
                continue
            if self.policy_eval is None:             # DummyNet (search/tests.rs:29-34): policy all ones, eval 0
                ones = np.ones(oracle.policy_size(n), dtype=np.float32)
                for i in pend:
                    self.nodes[i].devirtualize(ones, 0.0)
            else:
                pol, val = self.policy_eval([self.nodes[i].pending_state(0) for i in pend])
                for k, i in enumerate(pend):
                    self.nodes[i].devirtualize(pol[k], float(val[k]))
        # ---- pick, record, play, finish (:212-258)
        self.moves_played = [None] * self.workers
        for i in live:
            game, node = self.games[i], self.nodes[i]
            st = game.state()
            mv, vis, _, _, _ = node.children()
            if st.ply >= self.exploit_plies or self.sample is None:
                my_move = node.pick_move()                                            # pick_move(true)
            else:
                my_move = self.sample(i, self.serial[i], list(zip(mv.tolist(), vis.tolist())))
            self.incomplete[i].append((st, list(zip(mv.tolist(), vis.tolist()))))     # improved_policy (:220-223)
            node.play(my_move)
            game.play(my_move)
            self.moves_played[i] = my_move
            result = game.result()
            if result != 0:
                self.completed_games += 1
                self._complete(i, result_to_number(result))
                self._restart(i)
