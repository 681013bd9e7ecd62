"""`train::pit` on the device engines (tak_b200/pit.py; train/src/pit.rs:15-96): PlayerBatch keeps, per game, exactly the
schedule of a single `Player`, and a pit between two engines behaves as the reference's does."""
import numpy as np
import pytest

import tak_b200 as tb
from tak_b200 import weights as W
from tak_b200.pit import PitResult, PlayerBatch, pit

pytestmark = pytest.mark.gpu


def _tree(eng, gid):
    mv, vis, pri, rew = eng.children(gid)
    return mv.tolist(), vis.tolist(), pri.view(np.uint32).tolist(), rew.view(np.uint32).tolist(), eng.root(gid)


def _engine(n, arch, games, seed):
    eng = tb.Engine(n, games, nodes_per_game=1 << 14, max_batch=256)
    eng.net_create(arch)
    eng.net_load_weights(W.random_weights(arch, seed=seed))
    return eng


def test_player_batch_equals_independent_players():
    n, batch = 5, 4
    a, b = _engine(n, 5, 3, 1), _engine(n, 5, 3, 1)
    gids = [0, 1, 2]
    for eng in (a, b):
        eng.reset(0, 3, 4)
        eng.play(gids, [tb.parse_move(m, n) for m in ("a1", "e1", "c3")])     # three different games
    pb = PlayerBatch(a, gids, batch)
    singles = [tb.Player(b, g, batch) for g in gids]
    for ply in range(4):
        searching = gids if ply % 2 == 0 else [0, 2]      # a game whose player is not to move gets no rollouts
        for _ in range(3):
            pb.rollout(searching)
            for g in searching:
                singles[g].rollout()
        for g in gids:
            assert _tree(a, g) == _tree(b, g)
        mv = pb.pick_move(gids)
        assert [int(m) for m in mv] == [singles[g].pick_move(True) for g in gids]
        pb.play_move(gids, [int(m) for m in mv])
        for g in gids:
            singles[g].play_move(int(mv[g]))
        for g in gids:
            assert _tree(a, g) == _tree(b, g)
            assert a.download([g])[0].key() == b.download([g])[0].key()
    a.close()
    b.close()


def test_pit_same_network_is_symmetric_and_deterministic():
    n, games = 5, 3
    new, old = _engine(n, 5, 2 * games, 7), _engine(n, 5, 2 * games, 7)
    log1, log2 = [], []
    r1 = pit(new, old, games=games, batch=4, rollouts=3, seed=5, max_plies=300, log=log1)
    r2 = pit(new, old, games=games, batch=4, rollouts=3, seed=5, max_plies=300, log=log2)
    assert (r1.wins, r1.losses, r1.draws) == (r2.wins, r2.losses, r2.draws) and sorted(log1) == sorted(log2)
    assert len(log1) == 2 * games == r1.wins + r1.losses + r1.draws
    # identical networks: the two games of an opening are the same game, so "new" wins one iff it loses the other
    by_game = {g: (res, ply) for g, res, ply in log1}
    for i in range(games):
        assert by_game[2 * i] == by_game[2 * i + 1]
    assert r1.wins == r1.losses
    new.close()
    old.close()


def test_pit_result_counts():
    r = PitResult()
    r.update(tb.RESULT_WHITE | tb.RESULT_FLAG, 0)
    r.update(tb.RESULT_WHITE, 1)
    r.update(tb.RESULT_DRAW, 0)
    r.update(tb.RESULT_ONGOING, 0)
    assert (r.wins, r.losses, r.draws) == (1, 1, 1) and r.win_rate() == 0.5
