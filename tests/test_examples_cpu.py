"""CPU tests of the replay-format row: the oracle's restatement of tak::Symmetry (tak/src/symm.rs) and
alpha_tak::Example (alpha-tak/src/example.rs) against the reference's own symmetry test, and the C ABI's HOST functions
(text format, move / state symmetries -- they need no GPU) against that oracle."""
import numpy as np
import pytest

import oracle
import tak_b200 as tb
from test_oracle_golden import _sym_move
from util import random_positions, splitmix, to_tb_state


def _policy_for(g, seed):
    """A synthetic improved policy: every legal move with a pseudo-random visit count (some zero)."""
    moves = g.possible_moves()
    return [(m, splitmix(seed * 977 + i) % 50) for i, m in enumerate(moves)]


def test_symmetrical_boards_with_oracle_symmetries(golden):
    # reference: tak/tests/symm.rs:3-27 -- the 8 images of a game stay images of each other under the 8 images of a move
    n = 5
    for seed in golden["symm_seeds"][:4]:
        g0 = oracle.Game(n)
        games = [oracle.Game.from_state(oracle.symmetry_game(g0, k)) for k in range(8)]
        while games[0].result() == 0:
            moves = games[0].possible_moves()
            mv = moves[seed % len(moves)]
            for k, g in enumerate(games):
                assert g.play(oracle.symmetry_move(mv, n, k)) == 0
            for k, g in enumerate(games):  # symmetries(game)[k] commutes with play
                assert bytes(g.state()) == bytes(oracle.symmetry_game(games[0], k))
        assert len({g.result() for g in games}) == 1


@pytest.mark.parametrize("n", [3, 5, 6, 8])
def test_symmetry_group_structure(n):
    # identity first, 4 rotations then 4 mirrored rotations (symm.rs:11-20); independent python restatement agrees
    g = random_positions(n, 1, seed=5, max_ply=30)[0]
    for mv in g.possible_moves():
        imgs = [oracle.symmetry_move(mv, n, k) for k in range(8)]
        assert imgs[0] == mv
        assert imgs == [_sym_move(mv, n, k) for k in range(8)]
        assert oracle.symmetry_move(imgs[1], n, 1) == imgs[2] and oracle.symmetry_move(imgs[2], n, 1) == imgs[3]
        assert oracle.symmetry_move(imgs[3], n, 1) == mv                       # rot^4 = id
        assert oracle.symmetry_move(imgs[4], n, 4) == mv                       # mirror^2 = id
        assert [tb.symmetry_move(mv, n, k) for k in range(8)] == imgs          # the C ABI host function
    for k in range(8):
        assert bytes(tb.symmetry_state(to_tb_state(g.state()), k)) == bytes(oracle.symmetry_game(g, k))


def test_to_tensors_properties():
    # example.rs:63-78: every symmetry's pi sums to 1 over the same multiset of values, row k is the repr of image k,
    # and the policy entry of a move's image equals the move's share of the visits
    for n in (5, 6):
        for gi, g in enumerate(random_positions(n, 4, seed=11, max_ply=40)):
            pol = _policy_for(g, gi)
            ex = oracle.Example(g, pol, -1.0)
            inputs, pi, z = ex.to_tensors()
            total = sum(v for _, v in pol)
            assert (z == -1.0).all()
            for k in range(8):
                gk = oracle.Game.from_state(oracle.symmetry_game(g, k))
                assert np.array_equal(inputs[k], gk.repr())
                assert abs(pi[k].sum() - 1.0) < 1e-5
                assert np.array_equal(np.sort(pi[k]), np.sort(pi[0]))
                for m, v in pol:
                    idx = oracle.move_index(oracle.symmetry_move(m, n, k), n)
                    assert pi[k][idx] == np.float32(v) / np.float32(total)


def test_example_text_round_trip_and_abi_agreement():
    # Display / FromStr (example.rs:81-133): the oracle and the C ABI write the same line and read it back identically
    for n in (5, 6):
        for gi, g in enumerate(random_positions(n, 6, seed=3, max_ply=60)):
            g.set_half_komi(4)
            pol = [(m, v) for m, v in _policy_for(g, gi)]
            for result in (1.0, 0.0, -1.0):
                ex = oracle.Example(g, pol, result)
                line = str(ex)
                rec = tb.ReplayRecord()
                rec.state = to_tb_state(g.state())
                rec.result = result
                rec.n_children = len(pol)
                for i, (m, v) in enumerate(pol):
                    rec.moves[i], rec.visits[i] = m, v
                assert tb.example_format(rec) == line
                # fields: tps;ws;wc;bs;bc;half_komi;result;policy
                f = line.split(";")
                assert len(f) == 8 and f[0] == g.tps() and f[5] == "4" and f[6] == str(int(result))
                back = oracle.Example.parse(line, n)
                assert back.policy == pol and back.result == result
                rec2 = tb.example_parse(line + "\n", n)
                assert rec2.n_children == len(pol) and rec2.result == result
                assert list(rec2.moves[:len(pol)]) == [m for m, _ in pol]
                assert list(rec2.visits[:len(pol)]) == [v for _, v in pol]
                # TPS drops komi / reversible plies; the explicit fields restore reserves and half_komi (example.rs:108-113)
                s1, s2 = back.game.state(), rec2.state
                assert bytes(s1) == bytes(s2)
                assert s2.half_komi == 4 and s2.white_stones == g.state().white_stones
                assert tb.example_format(rec2) == line
    with pytest.raises(tb.TakNativeError):
        tb.example_parse("x5/x5/x5/x5/x5 1 1;21;1;21;1;0;1", 5)          # missing policy
    with pytest.raises(tb.TakNativeError):
        tb.example_parse("x5/x5/x5/x5/x5 1 1;21;1;21;1;0;1;a1-3", 5)     # pair has missing delimiter
