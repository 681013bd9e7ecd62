"""Full-size GPU checks through size-independent properties (the oracle cannot run these sizes in seconds):
BASELINE.json configs[2] (6x6, thousands of concurrent games, 800 rollouts/move, Net6) and configs[4] (8x8 stress:
deep stacks, long spreads, DummyNet MCTS)."""
import numpy as np
import pytest

import oracle
import tak_b200 as tb
from tak_b200 import weights as W
from util import random_positions, to_tb_state

pytestmark = pytest.mark.gpu


def test_search_6x6_800_rollouts_thousands_of_games():
    """2 046 concurrent 6x6 searches x 800 rollouts x Net6 (the bench's e2e path: upload -> mcts_rollouts ->
    mcts_children_batch -> mcts_pick_move).  Properties:
    * conservation: every root has 800 visits and its children's visits sum to 799 (the first rollout only expands,
      mcts.rs:36-47);
    * batch-position independence at full size: all games with the same forced opening (game id parity,
      self_play.rs:113-114) hold the SAME position, so their 800-rollout trees must be bit-identical although their
      leaves sit in 1 023 different positions of every evaluation batch -> exactly 2 distinct (moves, visits) vectors;
    * determinism: a second engine running the same schedule reproduces them;
    * the picked move is the LAST child with the maximal visit count (play.rs:52-58) and is legal."""
    G, R = 2046, 800
    ids = np.arange(G, dtype=np.int32)
    a1, a6, f6 = (tb.parse_move(m, 6) for m in ("a1", "a6", "f6"))
    results = []
    for run in range(2):
        eng = tb.Engine(6, G, nodes_per_game=1 << 17, max_batch=G)
        eng.net_create(6)
        eng.net_load_weights(W.random_weights(6, seed=5))
        eng.reset(0, G, 4)
        assert not eng.play(ids, [a1] * G).any()
        assert not eng.play(ids, [a6 if gid & 1 else f6 for gid in range(G)]).any()
        eng.tree_reset(ids)
        eng.rollouts(ids, R)
        mv, vis, cnt = eng.children_batch(ids, 256)
        picks = eng.pick_move(ids)
        for gid in (0, 1, G - 2, G - 1):
            assert eng.root(gid)[0] == R
        sums = vis.astype(np.int64).sum(axis=1)
        assert (sums == R - 1).all(), (sums.min(), sums.max())
        sig = {}
        for gid in range(G):
            k = int(cnt[gid])
            sig.setdefault((gid & 1, mv[gid, :k].tobytes(), vis[gid, :k].tobytes()), []).append(gid)
            best = int(mv[gid, k - 1 - int(np.argmax(vis[gid, :k][::-1]))])
            assert int(picks[gid]) == best
        assert len(sig) == 2, f"{len(sig)} distinct search results for 2 distinct positions"
        legal = eng.possible_moves([0, 1])
        assert int(picks[0]) in legal[0] and int(picks[1]) in legal[1]
        results.append(sorted(sig))
        eng.close()
    assert results[0] == results[1], "two runs of the same schedule differ"


def test_net6_forward_checksum_of_checksums_full_batch():
    """Linearity-free but size-independent: evaluating 5 328 boards (one full wave of conv tiles on 148 SMs, 6 tiles per
    SM, groups of 3) must give, board for board, the bits of evaluating them 37 at a time (ragged last tile, one CTA
    group) -- the fused tower may not depend on tile grouping, CTA assignment or batch position."""
    G = 5328
    games = random_positions(6, 48, seed=21, max_ply=60)
    eng = tb.Engine(6, 64, nodes_per_game=64, max_batch=G)
    eng.net_create(6)
    eng.net_load_weights(W.random_weights(6, seed=9))
    states = [to_tb_state(games[i % len(games)].state()) for i in range(G)]
    pol_big, val_big = eng.policy_eval(states)
    for lo in (0, 37 * 71, G - 37):
        pol, val = eng.policy_eval(states[lo:lo + 37])
        assert np.array_equal(pol.view(np.uint32), pol_big[lo:lo + 37].view(np.uint32))
        assert np.array_equal(val.view(np.uint32), val_big[lo:lo + 37].view(np.uint32))
    # identical positions -> identical outputs anywhere in the batch
    for i in range(len(games)):
        same = list(range(i, G, len(games)))
        assert (pol_big[same].view(np.uint32) == pol_big[i].view(np.uint32)).all()
        assert (val_big[same].view(np.uint32) == val_big[i].view(np.uint32)).all()
    assert np.allclose(pol_big.sum(axis=1), 1.0, atol=1e-4)
    eng.close()


def test_stress_8x8_deep_positions_mcts_and_playouts():
    """configs[4]: 8x8 positions cut from long random playouts (ply 60-200: tall stacks, spreads of up to 8 pieces,
    u128 stack columns).  Move lists, DummyNet search trees (visits / priors / rewards bit-exact) and a further 40
    plies of lock-step random play are compared with the oracle."""
    n = 8
    games = random_positions(n, 24, seed=77, half_komi=4, min_ply=60, max_ply=200)
    assert max(max(g.state().height) for g in games) >= 6, "sample holds no deep stack"
    ids = list(range(len(games)))
    eng = tb.Engine(n, len(games), nodes_per_game=1 << 17)
    eng.net_create(0)
    eng.upload(ids, [to_tb_state(g.state()) for g in games])
    for gid, mv in enumerate(eng.possible_moves(ids)):
        assert list(mv) == games[gid].possible_moves()
    eng.tree_reset(ids)
    eng.rollouts(ids, 150)
    for gid, g in enumerate(games):
        s = oracle.Search(n)
        s.rollouts_dummy(g, 150)
        mv, vis, pri, rew = eng.children(gid)
        omv, ovis, opri, orew, _ = s.children()
        assert np.array_equal(mv, omv) and np.array_equal(vis, ovis)
        assert np.array_equal(pri.view(np.uint32), opri.view(np.uint32))
        assert np.array_equal(rew.view(np.uint32), orew.view(np.uint32))
    live = ids
    for ply in range(40):
        lists = eng.possible_moves(live)
        picks = []
        for k, gid in enumerate(live):
            want = games[gid].possible_moves()
            assert list(lists[k]) == want
            picks.append(want[(gid * 31 + ply * 17) % len(want)])
        assert not eng.play(live, picks).any()
        for gid, mv in zip(live, picks):
            games[gid].play(mv)
        res = eng.result(live)
        for k, gid in enumerate(live):
            assert int(res[k]) == games[gid].result()
        for s, gid in zip(eng.download(live), live):
            assert s.key() == bytes(games[gid].state())
        live = [gid for gid in live if games[gid].result() == 0]
        if not live:
            break
    eng.close()
