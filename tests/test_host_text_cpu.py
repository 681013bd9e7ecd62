"""The PTN / TPS surface of the boundary (tak::Game <-> takparse::Tps, Move FromStr / Display: tak/src/tps.rs:7-96) on the
host side of the C ABI, for every board size, against the oracle: TPS of deep random positions (tall stacks, walls, caps),
round trips, the symmetry images of those positions, and malformed input (status codes, never a crash)."""
import pytest

import oracle
import tak_b200 as tb
from util import random_positions


@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 8])
def test_tps_of_deep_positions_matches_oracle(n):
    games = random_positions(n, 24, seed=100 + n, half_komi=0, min_ply=4, max_ply={3: 12, 4: 30}.get(n, 150))
    tallest = 0
    for g in games:
        st = tb.TakState.from_buffer_copy(bytes(g.state()))
        tallest = max(tallest, max(st.height[: n * n]))
        text = tb.tps_format(st)
        assert text == g.tps()
        back = tb.tps_parse(n, text)
        # komi and the reversible-ply counter are not part of a TPS (tps.rs:37-96): compare with the oracle's own parse
        assert back.key() == bytes(oracle.Game.from_tps(n, text).state())
        assert tb.tps_format(back) == text
        for k in range(8):                                    # the 8 images of the position print like the oracle's
            img = tb.symmetry_state(st, k)
            assert tb.tps_format(img) == oracle.Game.from_state(oracle.symmetry_game(g, k)).tps()
    assert tallest >= {3: 2, 4: 3}.get(n, 5)                   # the sample really contains stacks


@pytest.mark.parametrize("n", [3, 5, 6, 8])
def test_every_legal_move_text_round_trips(n):
    for g in random_positions(n, 6, seed=7 * n, min_ply=6, max_ply=60):
        for mv in g.possible_moves():
            text = oracle.format_move(mv, n)
            assert tb.format_move(mv, n) == text and tb.parse_move(text, n) == mv
            for k in range(8):
                assert tb.symmetry_move(mv, n, k) == oracle.symmetry_move(mv, n, k)


@pytest.mark.parametrize("text", ["", "a", "a0", "z1", "Xa1", "a1>", "3a1", "3a1>", "3a1>4", "9a1>9", "a1+0", "1a1+11",
                                  "Sa1+", "a1>>", "3a1x12", "  ", "a1 b2", "Ca1", "a1'", "a1!", "3c3>12", "e5", "f1",
                                  "5e5<1112", "6a1>", "2b2-11*"])
def test_move_text_is_accepted_or_rejected_like_the_oracle(text):
    """Same language on both sides: a string parses to the same move, or is refused with TAK_ERR_PARSE (never a crash)."""
    try:
        want = oracle.parse_move(text, 5)
    except Exception:
        want = None
    if want is None:
        with pytest.raises(tb.TakNativeError) as ex:
            tb.parse_move(text, 5)
        assert ex.value.code == -35            # TAK_ERR_PARSE
    else:
        assert tb.parse_move(text, 5) == want


@pytest.mark.parametrize("text", ["", "x5", "x5/x5/x5/x5/x5", "x5/x5/x5/x5/x5 1", "x5/x5/x5/x5/x4 1 1", "x5/x5/x5/x5/x6 1 1",
                                  "x5/x5/x5/x5/x5/x5 1 1", "x5/x5/x5/x5/x5 3 1", "x5/x5/x5/x5/x5 1 0",
                                  "1,2,x3/x5/x5/x5/x5 1 1 1", "3S,x4/x5/x5/x5/x5 1 1", "12Z,x4/x5/x5/x5/x5 1 1",
                                  "x5/x5/x5/x5/x5 one 1"])
def test_malformed_tps_is_rejected(text):
    with pytest.raises(tb.TakNativeError) as ex:
        tb.tps_parse(5, text)
    assert ex.value.code == -35            # TAK_ERR_PARSE


@pytest.mark.parametrize("text", ["x5/x5/x5/x5/x5 1 1", "1,2,x3/x5/x5/x5/x5 2 7", "x5/x5/x5/x5/x5 1 1 ",
                                  "x2,1,x2/x5/x5/x5/x5 1 2", "12S,x4/x5/x5/x5/x5 1 3", "2121C,x4/x5/x5/x5/x5 2 12"])
def test_well_formed_tps_matches_oracle(text):
    assert tb.tps_parse(5, text).key() == bytes(oracle.Game.from_tps(5, text).state())
