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


def _splitmix(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


@pytest.mark.parametrize("n,arch,G,R,passes", [(3, 0, 12, 40, 40), (4, 0, 12, 40, 40), (5, 5, 6, 32, 8), (6, 6, 6, 32, 8)])
def test_selfplay_with_instant_win_matches_reference_loop(n, arch, G, R, passes):
    """The bench's configuration minus the two thread_rng consumers: instant-win shortcut ON, always exploit, noise off.
    The device loop is compared pass by pass with oracle/selfplay_ref.py (the restatement of self_play.rs:96-262): every
    slot's position after each pass, and every completed Example (position, (move, visits) list incl. the 1000/1 fake
    visits of an instant win, result from the mover's perspective) -- including the reference's quirk that a slot reset
    by an instant win is searched from the empty board in the same pass."""
    from oracle.selfplay_ref import SelfPlayParallel
    seed, base = 0x7A4B, 1000
    eng = _engine(n, arch, G)
    eng.selfplay_begin(rollouts=R, half_komi=4, instant_win=1, exploit_ply=0, noise_ply=0, seed=seed, game_id_base=base)

    def coin(slot, serial):   # k_sp_opening's draw: a<N> when the hash is odd
        return bool(_splitmix(seed ^ _splitmix(((base + slot) << 32) | serial)) & 1)

    def policy_eval(states):
        return eng.policy_eval([to_tb_state(s) for s in states])

    ref = SelfPlayParallel(n, G, R, policy_eval=policy_eval if arch else None, coin=coin, exploit_plies=0,
                           instant_win=True)
    ids = list(range(G))
    got = []
    instant = 0
    for it in range(passes):
        st = eng.selfplay_step(1)
        assert st.records_truncated == 0
        got += eng.selfplay_drain(1 << 14)
        ref.iteration()
        after = eng.download(ids)
        for gid in ids:
            assert after[gid].key() == bytes(ref.games[gid].state()), f"pass {it} slot {gid}: positions differ"
    want = {(base + e.slot, e.serial, int(e.state.ply)): e for e in ref.examples}
    seen = set()
    for rec in got:
        key = (rec.game_id, rec.game_serial, int(rec.state.ply))
        assert key in want and key not in seen, key
        seen.add(key)
        e = want[key]
        assert bytes(rec.state) == bytes(e.state)
        assert [(rec.moves[i], rec.visits[i]) for i in range(rec.n_children)] == e.policy, key
        assert rec.result == e.result
        instant += any(v == 1000 for _, v in e.policy) and all(v in (1, 1000) for _, v in e.policy)
    assert seen == set(want), "completed examples differ"
    if arch == 0:
        assert instant >= 3 and ref.completed_games >= G, (instant, ref.completed_games)
    eng.close()
