"""GPU tests of the device self-play loop (reference: train/src/self_play.rs:96-262)."""
import math

import numpy as np
import pytest

import oracle
import tak_b200 as tb
from tak_b200 import weights as W
from util import to_tb_state

pytestmark = pytest.mark.gpu


def _engine(n, arch, games, rollouts_cap=1 << 15):
    eng = tb.Engine(n, games, nodes_per_game=rollouts_cap, max_batch=games)
    eng.net_create(arch)
    if arch:
        eng.net_load_weights(W.random_weights(arch, seed=2))
    return eng


@pytest.mark.parametrize("n,arch", [(5, 5), (6, 6)])
def test_selfplay_matches_oracle_loop(n, arch):
    """Noise off, always exploit, no instant-win shortcut: every ply the device loop plays must be the move an
    oracle-driven self_play_parallel (same schedule: one leaf per tree per step, tree reuse) picks when it is fed the
    engine's own policy_eval outputs; positions after each ply are compared bit for bit."""
    G, R, plies = 6, 48, 10
    eng = _engine(n, arch, G)
    eng.selfplay_begin(rollouts=R, half_komi=4, instant_win=0, exploit_ply=0, noise_ply=0, seed=7)
    ids = list(range(G))
    games = [None] * G
    searches = [oracle.Search(n) for _ in range(G)]
    for ply in range(plies):
        st = eng.selfplay_step(1)
        assert st.plies_played == G and st.rollouts == G * R
        eng.selfplay_drain(4096)
        after = eng.download(ids)
        for gid in ids:
            if games[gid] is None:
                # forced opening a1 + a<N>|<last><N> (self_play.rs:110-116): adopt the engine's coin flip
                a, b = oracle.Game(n, 4), oracle.Game(n, 4)
                a.play("a1"); a.play(f"a{n}")
                b.play("a1"); b.play(f"{'abcdefgh'[n - 1]}{n}")
                games[gid] = (a, b)
            cands = games[gid] if isinstance(games[gid], tuple) else (games[gid],)
            ok = None
            for g in cands:
                s = oracle.Search(n) if isinstance(games[gid], tuple) else searches[gid]
                g = g.clone()
                for _ in range(R):
                    if s.virtual_rollout(g) == 0:
                        pol, val = eng.policy_eval([to_tb_state(s.pending_state(0))])
                        s.devirtualize(pol[0], float(val[0]))
                mv = s.pick_move()
                g.play(mv)
                if g.state().key() == after[gid].key() or (g.result() != 0 and after[gid].ply == 0):
                    s.play(mv)
                    ok = (g, s)
                    break
            assert ok is not None, f"game {gid} ply {ply}: device move differs from the oracle loop"
            games[gid], searches[gid] = ok
            if games[gid].result() != 0:
                games[gid], searches[gid] = None, oracle.Search(n)
    eng.close()


def test_selfplay_records_and_restart():
    """Dummy network, instant-win shortcut on, sampling + noise on: games finish, slots restart, replay records are
    well-formed Examples (example.rs:28-33): legal move list in movegen order, visit counts, result in {1,0,-1}."""
    n, G, R = 4, 64, 24
    eng = _engine(n, 0, G, rollouts_cap=1 << 13)
    eng.selfplay_begin(rollouts=R, half_komi=4, instant_win=1, exploit_ply=6, noise_ply=8, seed=3, max_plies=20)
    total_games, records = 0, []
    for _ in range(14):
        st = eng.selfplay_step(2)
        total_games += st.games_completed
        records += eng.selfplay_drain(8192)
        assert st.kernel_launches > 0 and st.device_ms > 0
    assert total_games > G // 2, total_games
    assert len(records) > total_games  # several plies per finished game
    seen_games = set()
    for rec in records:
        assert rec.result in (1.0, 0.0, -1.0) and not math.isnan(rec.result)
        g = oracle.Game.from_state(oracle.TakState.from_buffer_copy(bytes(rec.state)))
        want = g.possible_moves()
        assert list(rec.moves[: rec.n_children]) == want
        vis = list(rec.visits[: rec.n_children])
        assert sum(vis) > 0
        seen_games.add((rec.game_id, rec.game_serial))
    assert len(seen_games) >= total_games - G  # the last step's games may still be held
    eng.close()
