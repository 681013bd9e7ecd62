"""Pins the CPU oracle against every known answer the reference's own tests hold for the hot path
(SURVEY.md section 8c): perft counts, win detection, golden TPS, TPS round trips, symmetric playouts,
board_repr planes, the 1575-entry 5x5 move table and the DummyNet MCTS behaviour tests."""
import numpy as np
import pytest

import oracle


def test_perft_known_answers(golden):
    # reference: tak/tests/perft.rs:20-99
    for case in golden["perft"]:
        g = oracle.Game.from_ptn_moves(case["n"], case["moves"])
        for depth, expect in case["expect"]:
            if expect > 3_500_000:
                continue  # the two d4 cases run in test_perft_deep (threads)
            assert g.perft(depth) == expect, (case["name"], depth)


def test_perft_deep(golden):
    # 5x5 d4 = 2 999 784, 6x6 d4 = 13 586 048, endgame d4 = 16 642 760 (perft.rs:64,88,96)
    for case in golden["perft"]:
        g = oracle.Game.from_ptn_moves(case["n"], case["moves"])
        for depth, expect in case["expect"]:
            if expect > 3_500_000:
                assert g.perft(depth, threads=8) == expect, (case["name"], depth)


def test_wins(golden):
    # reference: tak/tests/wins.rs:5-67
    for case in golden["wins"]:
        g = oracle.Game.from_ptn_moves(case["n"], case["moves"])
        for chk in case["checks"]:
            if chk["half_komi"] is not None:
                g.set_half_komi(chk["half_komi"])
            assert g.result() == chk["result"], case["name"]


def test_golden_tps(golden):
    # reference: tak/tests/tps.rs:5-24
    t = golden["tps"]
    g = oracle.Game.from_ptn_moves(t["n"], t["moves"])
    assert g.tps() == t["tps"]


def _playout(n, seed, visit):
    g = oracle.Game(n)
    plies = 0
    while g.result() == 0:
        moves = g.possible_moves()
        mv = moves[seed % len(moves)]
        assert g.play(mv) == 0
        plies += 1
        visit(g, mv)
    return g, plies


def test_tps_consistency(golden):
    # reference: tak/tests/tps.rs:26-96
    t = golden["tps"]
    for seed in t["consistency_seeds"]:
        def check(g, _mv):
            s = g.state()
            s2 = oracle.Game.from_tps(g.n, g.tps()).state()
            assert bytes(s.height) == bytes(s2.height)
            assert bytes(s.top) == bytes(s2.top)
            assert bytes(s.stack_lo) == bytes(s2.stack_lo)
            for f in ("to_move", "ply", "white_caps", "white_stones", "black_caps", "black_stones"):
                assert getattr(s, f) == getattr(s2, f), f
        _playout(t["consistency_n"], seed, check)


def test_playout_lengths():
    # deterministic moves[seed % len] playouts on 5x5 end after 77 / 101 / 165 plies
    for seed, plies in ((5915587277, 77), (1500450271, 101), (3267000013, 165)):
        _, k = _playout(5, seed, lambda g, m: None)
        assert k == plies


def _sym_square(col, row, n, k):
    # k-th of the 8 symmetries: rotations then mirrored rotations (any consistent choice is a symmetry)
    c, r = col, row
    if k >= 4:
        c = n - 1 - c
    for _ in range(k % 4):
        c, r = r, n - 1 - c
    return c, r


def _sym_dir(d, k):
    # directions: 0 Up 1 Down 2 Left 3 Right ; as unit vectors
    vec = {0: (0, 1), 1: (0, -1), 2: (-1, 0), 3: (1, 0)}[d]
    dc, dr = vec
    if k >= 4:
        dc = -dc
    for _ in range(k % 4):
        dc, dr = dr, -dc
    return {(0, 1): 0, (0, -1): 1, (-1, 0): 2, (1, 0): 3}[(dc, dr)]


def _sym_move(mv, n, k):
    sq = mv & 63
    col, row = sq % n, sq // n
    c, r = _sym_square(col, row, n, k)
    out = (mv & 0xFF00) | (r * n + c)
    kind = (mv >> 6) & 3
    if mv >> 8:
        kind = _sym_dir(kind, k)
    return out | (kind << 6)


def test_symmetrical_boards(golden):
    # reference: tak/tests/symm.rs:3-27 -- symmetric games played with symmetric moves end equally
    n = 5
    for seed in golden["symm_seeds"]:
        games = [oracle.Game(n) for _ in range(8)]
        while games[0].result() == 0:
            moves = games[0].possible_moves()
            mv = moves[seed % len(moves)]
            for k, g in enumerate(games):
                assert g.play(_sym_move(mv, n, k)) == 0
        res = [g.result() for g in games]
        assert len(set(res)) == 1


def test_board_repr_golden(golden):
    # reference: alpha-tak/src/repr/tests.rs:11-111
    assert not oracle.Game(5).board_repr(0).any()
    r = golden["board_repr"]
    g = oracle.Game.from_ptn_moves(r["n"], r["moves"])
    planes = g.board_repr(r["to_move_arg"])
    want = np.zeros_like(planes)
    want[:12] = np.array(r["planes_12x5x5"], dtype=np.float32).reshape(12, 5, 5)
    assert np.array_equal(planes, want)


def test_game_repr_shape_and_scalars():
    g = oracle.Game.from_ptn_moves(6, ["a1", "f6", "c3", "Sd4", "Cc4"], half_komi=4)
    x = g.repr()
    assert x.shape == (92, 6, 6) and oracle.input_channels(5) == 72 and oracle.input_channels(8) == 138
    # black to move: colour plane 0; fcd = (white flats - black flats) - half_komi/2 = (2 - 1) - 2 = -1
    assert np.all(x[90] == 0.0) and np.allclose(x[91], -1 / 36)
    # board_channels(6) = 28; reserves one-hot: mover (black) has 28 stones -> plane 28+27, white 28 -> 28+30+27
    assert x[28 + 27].all() and x[28 + 30 + 27].all() and x[28:88].sum() == 72
    assert x[88].all() and not x[89].any()  # mover (black) still has its cap, white has played it


def test_move_index_5_table(golden_moves_5):
    # reference: alpha-tak/src/search/move_map.rs:51-201 -- the oracle regenerates the list by rule
    assert oracle.legacy_moves_5() == golden_moves_5
    for i, text in enumerate(golden_moves_5):
        assert oracle.move_index(oracle.parse_move(text, 5), 5) == i
        assert oracle.format_move(oracle.parse_move(text, 5), 5) == text


def test_move_index_6_formula():
    # reference: move_map.rs:26-47 ; "3a1>21" -> mask 0b0110_0000 -> (>>2)-1 = 23, Right = 1
    assert oracle.policy_size(6) == 9036 and oracle.policy_size(5) == 1575 and oracle.policy_size(8) == 65216
    mv = oracle.parse_move("3a1>21", 6)
    assert mv >> 8 == 0b01100000
    assert oracle.move_index(mv, 6) == (3 + 23 + 62 * 1) * 36 + 0
    assert oracle.move_index(oracle.parse_move("Cc2", 6), 6) == 2 * 36 + 1 * 6 + 2
    # every legal move of a busy position maps to a distinct in-range index
    g = oracle.Game.from_ptn_moves(6, ["a1", "f6", "c3", "d4", "c4", "d3", "c4-", "d3<", "Cb2", "d5"])
    idx = [oracle.move_index(m, 6) for m in g.possible_moves()]
    assert len(set(idx)) == len(idx) and min(idx) >= 0 and max(idx) < 9036


def test_mcts_dummy_net(golden):
    # reference: alpha-tak/src/search/tests.rs:38-72 (DummyNet: policy 1.0, eval 0.0)
    cases = {c["name"]: c for c in golden["mcts"]}
    c = cases["win_in_one"]
    g = oracle.Game.from_ptn_moves(c["n"], c["moves"])
    s = oracle.Search(c["n"])
    s.rollouts_dummy(g, c["rollouts"])
    g.play(s.pick_move())
    assert g.result() == 0x11  # white road

    c = cases["prevent_win_in_two"]
    g = oracle.Game.from_ptn_moves(c["n"], c["moves"])
    s = oracle.Search(c["n"])
    s.rollouts_dummy(g, c["rollouts"])
    mv = s.pick_move()
    s.play(mv)
    g.play(mv)
    assert g.result() == 0
    s.rollouts_dummy(g, c["rollouts"])
    g.play(s.pick_move())
    assert g.result() == 0


def test_mcts_stepwise_equals_fused():
    # virtual_rollout + devirtualize (what the GPU path mirrors) == Node::rollout
    n = 5
    g = oracle.Game.from_ptn_moves(n, ["a1", "e5", "c3", "c2", "d3"])
    a, b = oracle.Search(n), oracle.Search(n)
    a.rollouts_dummy(g, 300)
    ones = np.ones(oracle.policy_size(n), dtype=np.float32)
    for _ in range(300):
        if b.virtual_rollout(g) == 0:
            b.devirtualize(ones, 0.0)
    for x, y in zip(a.children(), b.children()):
        assert np.array_equal(x, y)
    assert a.root() == b.root()
