"""GPU parity of the tak::Game path (movegen order, play, result, perft) against the CPU oracle and the
reference's known answers.  Everything goes through the C ABI (tak_b200 -> libtaknative.so)."""
import numpy as np
import pytest

import oracle
import tak_b200 as tb
from util import random_positions, to_tb_state

pytestmark = pytest.mark.gpu


def splitmix(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def test_perft_known_answers(golden):
    # reference: tak/tests/perft.rs:20-99 -- read like the reference's own test
    for case in golden["perft"]:
        game = tb.Game.from_ptn_moves(case["n"], case["moves"])
        for depth, expect in case["expect"]:
            assert game.perft(depth) == expect, (case["name"], depth)


def test_perft_multi_shares_a_frontier(golden):
    """tak_perft_multi (multi-GPU perft): perft(root, d) == lines that ended above the cut +
    sum over ranks of perft_multi(every world-th position of the depth-k frontier, d - k) -- on the reference's
    positional cases (endgame_perft ends lines early) and on the openings."""
    eng = tb.Engine(5, 4096, nodes_per_game=64)
    eng6 = tb.Engine(6, 2048, nodes_per_game=64)
    for case in golden["perft"]:
        e = eng if case["n"] == 5 else eng6
        game = tb.Game.from_ptn_moves(case["n"], case["moves"])
        depth, expect = max((d, x) for d, x in case["expect"] if x < 5_000_000)
        for cut in (1, 2):
            if cut > depth:
                continue
            front, ended = e.frontier(game.state(), cut)
            for world in (1, 3):
                total = ended + sum(e.perft_multi(front[r::world], depth - cut) for r in range(world))
                assert total == expect, (case["name"], depth, cut, world)
    assert eng.perft_multi([], 3) == 0
    eng.close()
    eng6.close()


def test_perft_6x6_depth5():
    # the value the reference keeps commented out (perft.rs:98)
    assert tb.Game.default(6).perft(5) == 1_253_506_520


def test_wins(golden):
    # reference: tak/tests/wins.rs:5-67
    for case in golden["wins"]:
        game = tb.Game.from_ptn_moves(case["n"], case["moves"])
        for chk in case["checks"]:
            if chk["half_komi"] is not None:
                game.set_half_komi(chk["half_komi"])
            assert game.result() == chk["result"], case["name"]


def test_golden_tps(golden):
    # reference: tak/tests/tps.rs:5-24
    t = golden["tps"]
    assert tb.Game.from_ptn_moves(t["n"], t["moves"]).tps() == t["tps"]


@pytest.mark.parametrize("n,games,komi", [(3, 64, 0), (4, 64, 1), (5, 256, 0), (6, 256, 4), (7, 32, 0), (8, 48, 4)])
def test_random_playouts_match_oracle(n, games, komi):
    """Lock-step random playouts: at every ply the GPU's move list (order included), the state after the move
    and the result must equal the oracle's, bit for bit."""
    eng = tb.Engine(n, games)
    eng.reset(0, games, komi)
    orc = [oracle.Game(n, komi) for _ in range(games)]
    live = list(range(games))
    ply = 0
    max_moves = 0
    while live and ply < 400:
        lists = eng.possible_moves(live)
        picks = []
        for gid, mv in zip(live, lists):
            want = orc[gid].possible_moves()
            assert list(mv) == want, f"move list differs: n={n} game={gid} ply={ply}"
            max_moves = max(max_moves, len(want))
            picks.append(want[splitmix(gid * 1000003 + ply) % len(want)])
        st = eng.play(live, picks)
        assert not st.any()
        for gid, mv in zip(live, picks):
            assert orc[gid].play(mv) == 0
        res = eng.result(live)
        states = eng.download(live)
        nxt = []
        for gid, r, s in zip(live, res, states):
            assert int(r) == orc[gid].result(), f"result differs: n={n} game={gid} ply={ply}"
            assert s.key() == orc[gid].state().key(), f"state differs: n={n} game={gid} ply={ply}"
            if r == 0:
                nxt.append(gid)
        live = nxt
        ply += 1
    assert max_moves > 3 * n
    eng.close()


def test_play_error_codes_fuzz():
    """Game::play on arbitrary (mostly illegal) moves: the status must be the PlayError the reference would raise
    (check order of tak/src/game.rs:147-209, tile.rs:28-63) and legal ones must produce the oracle's state."""
    seen = set()
    for n in (4, 5, 6, 8):
        games = random_positions(n, 48, seed=40 + n, max_ply=70)
        eng = tb.Engine(n, len(games))
        ids = list(range(len(games)))
        states = [to_tb_state(g.state()) for g in games]
        for rnd in range(24):
            eng.upload(ids, states)
            moves = []
            for gid in ids:
                r = splitmix(gid * 7919 + rnd * 104729 + n)
                sq = r % (n * n)
                kind = (r >> 8) % 4
                mask = (r >> 16) & 0xFF if (r >> 12) % 4 else 0
                if mask == 0 and kind == 3:
                    kind = 0
                if (r >> 40) % 3 == 0:  # bias towards the mover's own stacks with plausible pickups
                    legal = [m for m in games[gid].possible_moves() if m >> 8]
                    if legal:
                        base = legal[(r >> 44) % len(legal)]
                        sq, kind = base & 63, (base >> 6) & 3
                moves.append(sq | (kind << 6) | (mask << 8))
            st = eng.play(ids, moves)
            after = eng.download(ids)
            for gid in ids:
                o = games[gid].clone()
                want = o.play(moves[gid])
                assert int(st[gid]) == want, (n, gid, hex(moves[gid]), int(st[gid]), want)
                seen.add(want)
                if want == 0:
                    assert after[gid].key() == o.state().key()
        eng.close()
    assert {0, -2, -6, -7, -12, -13}.issubset(seen), seen


def test_play_error_codes_opening():
    g = tb.Game.default(5)
    assert g.play("Sa1") == -5 and tb.Game.default(5).play("Ca1") == -5   # OpeningNonFlat
    g = tb.Game.from_ptn_moves(5, ["a1"])
    assert g.play("a1") == -2                                             # AlreadyOccupied
    g = tb.Game.from_ptn_moves(5, ["a1", "b1", "Cc1", "d1"])
    assert g.play("Ce1") == -3                                            # NoCapstone


def test_ptn_tps_text_matches_oracle(golden_moves_5):
    for i, text in enumerate(golden_moves_5):
        mv = tb.parse_move(text, 5)
        assert mv == oracle.parse_move(text, 5)
        assert tb.format_move(mv, 5) == text
        assert tb.move_index(mv, 5) == i
    g = random_positions(6, 1, seed=77, min_ply=30, max_ply=60)[0]
    for mv in g.possible_moves():
        assert tb.move_index(mv, 6) == oracle.move_index(mv, 6)
    s = tb.tps_parse(6, g.tps())
    assert tb.tps_format(s) == g.tps()
    assert s.key() == oracle.Game.from_tps(6, g.tps()).state().key()


def test_game_surface_helpers():
    """The rest of the `tak::Game` / `Board` surface: safe_play, flat_diff, full, symmetries, default_starting_stones."""
    import oracle
    from util import random_positions
    assert tb.default_starting_stones(6) == (30, 1) and tb.default_starting_stones(8) == (50, 2)
    for og in random_positions(5, 6, seed=9, max_ply=40):
        g = tb.Game.from_state(tb.TakState.from_buffer_copy(bytes(og.state())))
        assert g.flat_diff() == og.flat_diff()
        before = g.state().key()
        mv = g.possible_moves()[0]
        backup = g.safe_play(mv)
        assert backup.state().key() == before and g.state().key() != before
        with pytest.raises(tb.TakNativeError):
            g.safe_play(tb.parse_move("a1", 5) if g.state().height[0] else tb.parse_move("1a1+", 5))
        og.play(mv)
        assert g.state().key() == bytes(og.state())           # a failed safe_play leaves the game untouched
        syms = g.symmetries()
        assert len(syms) == 8 and syms[0].state().key() == g.state().key()
        assert sorted(s.flat_diff() for s in syms) == [g.flat_diff()] * 8
        assert all(s.board_full() == g.board_full() for s in syms)


def test_symmetrical_boards(golden):
    # reference: tak/tests/symm.rs:3-68 -- read like the reference's own test, on the device engine
    eng = tb.Engine(5, 8, nodes_per_game=64)
    ids = list(range(8))
    for seed in golden["symm_seeds"]:
        eng.reset(0, 8, 0)
        start = eng.download([0])[0]
        eng.upload(ids, [tb.symmetry_state(start, k) for k in range(8)])       # Game::<5>::default().symmetries()
        plies = 0
        while (int(eng.result([0])[0]) & 3) == tb.RESULT_ONGOING:
            moves = eng.possible_moves([0])[0]
            my_move = int(moves[seed % len(moves)])
            status = eng.play(ids, [tb.symmetry_move(my_move, 5, k) for k in range(8)])
            assert not status.any(), (seed, plies, status)
            plies += 1
        res = [int(r) for r in eng.result(ids)]
        assert len(set(res)) == 1, (seed, res)
        # stronger than the reference: the 8 games are still each other's images
        states = eng.download(ids)
        assert [s.key() for s in states] == [tb.symmetry_state(states[0], k).key() for k in range(8)]
    eng.close()


def test_tps_consistency(golden):
    # reference: tak/tests/tps.rs:26-96 -- after every ply of a deterministic playout, Game -> Tps -> Game keeps the board,
    # the side to move, the ply and the four reserve counters (komi and the reversible-ply counter are not in a TPS)
    for seed in golden["symm_seeds"]:
        game = tb.Game.default(5)
        while (game.result() & 3) == tb.RESULT_ONGOING:
            moves = game.possible_moves()
            assert game.play(moves[seed % len(moves)]) == 0
            st = game.state()
            back = tb.tps_parse(5, tb.tps_format(st))
            nsq = 25
            assert list(back.height[:nsq]) == list(st.height[:nsq]) and list(back.top[:nsq]) == list(st.top[:nsq])
            assert list(back.stack_lo[:nsq]) == list(st.stack_lo[:nsq])
            assert (back.to_move, back.ply) == (st.to_move, st.ply)
            assert (back.white_caps, back.white_stones, back.black_caps, back.black_stones) == \
                   (st.white_caps, st.white_stones, st.black_caps, st.black_stones)
