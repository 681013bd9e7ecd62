"""GPU parity of the alpha_tak::Node path: bit-exact visit counts / rewards / priors against the CPU oracle when
both consume the same network outputs (reference semantics: alpha-tak/src/search/mcts.rs, play.rs)."""
import numpy as np
import pytest

import oracle
import tak_b200 as tb
from tak_b200 import weights as W
from util import random_positions, to_tb_state

pytestmark = pytest.mark.gpu


def assert_tree_equal(eng, gid, search):
    mv, vis, pri, rew = eng.children(gid)
    omv, ovis, opri, orew, _ = search.children()
    assert np.array_equal(mv, omv), "children moves/order differ"
    assert np.array_equal(vis, ovis), f"visit counts differ: {vis} vs {ovis}"
    assert np.array_equal(pri.view(np.uint32), opri.view(np.uint32)), "priors differ (bits)"
    assert np.array_equal(rew.view(np.uint32), orew.view(np.uint32)), "expected rewards differ (bits)"
    v, vv, r = eng.root(gid)
    ov, ovv, orr = search.root()
    assert (v, vv) == (ov, ovv)
    assert np.float32(r).view(np.uint32) == np.float32(orr).view(np.uint32)


def test_reference_mcts_tests_dummy_net(golden):
    # reference: alpha-tak/src/search/tests.rs:38-72, run through the device search
    cases = {c["name"]: c for c in golden["mcts"]}
    eng = tb.Engine(3, 2, nodes_per_game=1 << 15)
    eng.net_create(0)
    c = cases["win_in_one"]
    game = tb.Game.from_ptn_moves(c["n"], c["moves"])
    eng.upload([0], [game.state()])
    eng.tree_reset([0])
    eng.rollouts([0], c["rollouts"])
    game.play(int(eng.pick_move([0])[0]))
    assert game.result() == 0x11

    c = cases["prevent_win_in_two"]
    game = tb.Game.from_ptn_moves(c["n"], c["moves"])
    eng.upload([1], [game.state()])
    eng.tree_reset([1])
    eng.rollouts([1], c["rollouts"])
    mv = int(eng.pick_move([1])[0])
    eng.tree_play([1], [mv])
    game.play(mv)
    eng.upload([1], [game.state()])
    assert game.result() == 0
    eng.rollouts([1], c["rollouts"])
    game.play(int(eng.pick_move([1])[0]))
    assert game.result() == 0
    eng.close()


@pytest.mark.parametrize("n,rollouts", [(3, 600), (5, 500), (6, 400), (8, 200)])
def test_dummy_net_bit_exact_vs_oracle(n, rollouts):
    games = random_positions(n, 6, seed=3 * n, max_ply=40)
    eng = tb.Engine(n, len(games), nodes_per_game=1 << 17)
    eng.net_create(0)
    ids = list(range(len(games)))
    eng.upload(ids, [to_tb_state(g.state()) for g in games])
    eng.tree_reset(ids)
    eng.rollouts(ids, rollouts)
    searches = []
    for gid, g in enumerate(games):
        s = oracle.Search(n)
        s.rollouts_dummy(g, rollouts)
        assert_tree_equal(eng, gid, s)
        searches.append(s)
    # tree reuse (Node::play) + more rollouts
    picks = eng.pick_move(ids)
    for gid, (g, s) in enumerate(zip(games, searches)):
        assert int(picks[gid]) == s.pick_move()
        s.play(int(picks[gid]))
        g.play(int(picks[gid]))
    eng.tree_play(ids, picks)
    st = eng.play(ids, picks)
    assert not st.any()
    live = [i for i in ids if games[i].result() == 0]
    eng.rollouts(live, rollouts // 2)
    for gid in live:
        searches[gid].rollouts_dummy(games[gid], rollouts // 2)
        assert_tree_equal(eng, gid, searches[gid])
    eng.close()


def _net_engine(n, games, arch):
    eng = tb.Engine(n, games, nodes_per_game=1 << 16, max_batch=max(games * 8, 64))
    eng.net_create(arch)
    eng.net_load_weights(W.random_weights(arch, seed=1))
    return eng


@pytest.mark.parametrize("arch", [5, 6])
def test_stepwise_bit_exact_with_network(arch):
    """virtual_rollout -> (network) -> devirtualize, one leaf per tree per step (self_play.rs:181-210): the oracle is
    fed the engine's own fp32 policy_eval outputs and must end with identical visit counts."""
    n = arch
    games = random_positions(n, 8, seed=17 + arch, max_ply=30)
    ids = list(range(len(games)))
    eng = _net_engine(n, len(games), arch)
    eng.upload(ids, [to_tb_state(g.state()) for g in games])
    eng.tree_reset(ids)
    searches = [oracle.Search(n) for _ in games]
    for step in range(120):
        eng.virtual_rollout(ids, 1)
        gids, leaves = eng.pending()
        want = []
        for gid, (g, s) in enumerate(zip(games, searches)):
            if s.virtual_rollout(g) == 0:
                want.append(gid)
        assert list(gids) == want
        for gid, leaf in zip(gids, leaves):
            assert leaf.key() == bytes(searches[gid].pending_state(0)), f"leaf state differs at step {step}"
        pol, val = eng.policy_eval(leaves)
        for i, gid in enumerate(gids):
            searches[gid].devirtualize(pol[i], float(val[i]))
        eng.devirtualize()          # priors gathered on the device from the same network
    for gid in ids:
        assert_tree_equal(eng, gid, searches[gid])
    eng.close()


def test_fused_rollouts_and_host_supplied_outputs():
    n, arch = 6, 6
    games = random_positions(n, 4, seed=99, max_ply=24)
    ids = list(range(len(games)))
    eng = _net_engine(n, len(games), arch)
    eng.upload(ids, [to_tb_state(g.state()) for g in games])
    eng.tree_reset(ids)
    eng.rollouts(ids, 200)  # mcts_rollouts: select -> encode -> net -> backup entirely on device
    searches = [oracle.Search(n) for _ in games]
    for _ in range(200):
        for g, s in zip(games, searches):
            if s.virtual_rollout(g) == 0:
                pol, val = eng.policy_eval([to_tb_state(s.pending_state(0))])
                s.devirtualize(pol[0], float(val[0]))
    for gid in ids:
        assert_tree_equal(eng, gid, searches[gid])
    # mcts_devirtualize_with: the caller supplies Network::policy_eval outputs (drop-in for a foreign network)
    eng.virtual_rollout(ids, 1)
    gids, leaves = eng.pending()
    rng = np.random.default_rng(0)
    pol = rng.random((len(gids), tb.policy_size(n)), dtype=np.float32)
    val = rng.uniform(-1, 1, len(gids)).astype(np.float32)
    eng.devirtualize_with(pol, val)
    for i, gid in enumerate(gids):
        assert searches[gid].virtual_rollout(games[gid]) == 0
        searches[gid].devirtualize(pol[i], float(val[i]))
    for gid in ids:
        assert_tree_equal(eng, gid, searches[gid])
    eng.close()


def test_player_batch_semantics():
    """k virtual rollouts per tree before one batched evaluation (Player, player.rs:78-110): selections inside a batch
    see each other's virtual losses; devirtualisation happens in queue order."""
    n, k = 5, 8
    games = random_positions(n, 3, seed=5, max_ply=20)
    ids = list(range(len(games)))
    eng = tb.Engine(n, len(games), nodes_per_game=1 << 16)
    eng.net_create(0)
    eng.upload(ids, [to_tb_state(g.state()) for g in games])
    eng.tree_reset(ids)
    searches = [oracle.Search(n) for _ in games]
    ones = np.ones(tb.policy_size(n), dtype=np.float32)
    for _ in range(40):
        eng.virtual_rollout(ids, k)
        for g, s in zip(games, searches):
            for _ in range(k):
                s.virtual_rollout(g)
        gids, _ = eng.pending(with_states=False)
        assert len(gids) == sum(s.pending() for s in searches)
        eng.devirtualize()
        for s in searches:
            while s.pending():
                s.devirtualize(ones, 0.0)
    for gid in ids:
        assert_tree_equal(eng, gid, searches[gid])
    eng.close()
