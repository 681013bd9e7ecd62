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


def test_perft_from_deep_positions_all_sizes():
    """perf_count (tak/tests/perft.rs:3-18) from positions cut out of long random playouts, N = 3..8: the expansion
    patches packed records one thread per child (RecordPlay) -- tall stacks, multi-drop spreads, flattening capstones,
    u128 columns on 7x7 / 8x8 -- and classifies them from the tail it builds; counts must equal the oracle's."""
    for n, count, depth, lo, hi in ((3, 6, 4, 2, 14), (4, 6, 3, 4, 40), (5, 6, 3, 10, 80), (6, 6, 3, 20, 120),
                                    (7, 4, 2, 40, 160), (8, 4, 2, 60, 200)):
        games = random_positions(n, count, seed=300 + n, min_ply=lo, max_ply=hi)
        eng = tb.Engine(n, 8, nodes_per_game=64)
        for g in games:
            st = to_tb_state(g.state())
            for d in range(1, depth + 1):
                assert eng.perft(st, d) == g.perft(d), (n, d, g.tps())
        roots = [to_tb_state(g.state()) for g in games]
        assert eng.perft_multi(roots, depth) == sum(g.perft(depth) for g in games)
        eng.close()


def test_device_playouts_match_oracle():
    """tak_playouts: uniform-random playouts on the device (state in registers from the first to the last ply) against
    the same loop over the oracle -- plies, results and final positions, with a full and a staggered ply budget."""
    seed = 0x51ED
    for n, G in ((5, 12), (6, 12), (8, 10)):
        eng = tb.Engine(n, G, nodes_per_game=64)
        for max_plies, spread, base in ((10_000, 0, 0), (30, 25, 1000)):
            eng.reset(0, G, 4)
            plies, res, tot = eng.playouts(0, G, seed, max_plies, spread, game_id_base=base)
            states = eng.download(list(range(G)))
            gen = 0
            for gid in range(G):
                g = oracle.Game(n, 4)
                budget = max_plies + (splitmix(seed ^ splitmix((base + gid) ^ 0xC0FFEE)) % spread if spread else 0)
                k = 0
                while g.result() == 0 and k < budget:
                    moves = g.possible_moves()
                    gen += len(moves)
                    ply = g.state().ply
                    g.play(moves[splitmix(seed ^ splitmix(((base + gid) << 32) | ply)) % len(moves)])
                    k += 1
                assert k == int(plies[gid]) and g.result() == int(res[gid]), (n, gid)
                assert states[gid].key() == bytes(g.state()), (n, gid)
            assert tot["plies"] == int(plies.sum()) and tot["generated"] == gen
        eng.close()


def test_8x8_stacks_taller_than_64():
    """u128 stack columns above bit 63 (reference edge case, SURVEY.md section 7: "stacks deeper than 64"): 8x8 positions
    built from TPS (legal piece counts) with (a) a 70-piece stack owned by the mover -- carries are cut out across the
    64-bit boundary -- once with a flat and once with a capstone on top, and (b) a 62-piece stack next to a 10-piece stack
    of the mover -- drops push it over the boundary.  Move lists, every child position, results, the children's own move
    lists and perft(2) are compared with the oracle."""
    n = 8
    alt = lambda k, first: "".join("12"[(i + first) % 2] for i in range(k))
    cases = [
        (f"x8/x8/x8/x8/x3,2S,x4/x8/1,x,2,x5/{alt(70, 1)},{alt(24, 0)},x,1S,x4 1 70", 70, 63),   # white flat on 70
        (f"x8/x8/x8/x8/x3,1S,x4/x8/2,x,1,x5/{alt(69, 1)}2C,{alt(24, 1)},x,2S,x4 2 70", 70, 63),  # black cap on 70
        (f"x8/x8/x8/x8/x3,2S,x4/x8/1,x,2,x5/{alt(62, 0)},{alt(10, 1)},x,1S,x4 1 60", 62, 70),   # 62 + up to 8 dropped
    ]
    for tps, tallest_parent, tallest_child_min in cases:
        og = oracle.Game.from_tps(n, tps)
        st = tb.tps_parse(n, tps)
        assert bytes(st) == bytes(og.state()), tps
        assert max(st.height) == tallest_parent and max(st.white_stones, st.black_stones) <= 50
        eng = tb.Engine(n, 512, nodes_per_game=64)
        eng.upload([0], [st])
        moves = og.possible_moves()
        assert list(eng.possible_moves([0])[0]) == moves
        k = len(moves)
        ids = list(range(k))
        eng.upload(ids, [st] * k)
        assert not eng.play(ids, moves).any()
        res = eng.result(ids)
        tallest = 0
        children = []
        for i, (child, mv) in enumerate(zip(eng.download(ids), moves)):
            g = og.clone()
            assert g.play(mv) == 0
            assert child.key() == bytes(g.state()), tb.format_move(mv, n)
            assert int(res[i]) == g.result()
            tallest = max(tallest, max(child.height))
            children.append(g)
        assert tallest >= tallest_child_min and any(any(c.state().stack_hi) for c in children)
        for depth in (1, 2):
            assert eng.perft(st, depth) == og.perft(depth)
        # second ply: children whose stacks sit across the boundary generate moves too
        lists = eng.possible_moves(ids)
        for i, g in enumerate(children):
            if g.result() == 0:
                assert list(lists[i]) == g.possible_moves()
        eng.close()


def test_perft_sliced_frontiers():
    """A frontier whose children exceed the per-level arena is cut into slices of parents that are expanded and recursed
    into one after the other (memory stays bounded for deep perfts).  TAK_PERFT_CAP = 20 000 states forces that path at
    several levels of 6x6 perft(4) / 5x5 perft(4) and of a deep 8x8 position; the counts must not change."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "import tak_b200 as tb\n"
        "from util import random_positions, to_tb_state\n"
        "assert tb.Game.default(5).perft(4) == 2999784\n"
        "g = random_positions(8, 1, seed=5, min_ply=80, max_ply=120)[0]\n"
        "e = tb.Engine(8, 4, nodes_per_game=64)\n"
        "assert e.perft(to_tb_state(g.state()), 3) == g.perft(3, threads=8)\n"
        "e6 = tb.Engine(6, 4, nodes_per_game=64)\n"
        "assert e6.perft(tb.state_init(6, 0), 4) == 13586048\n"
        "print('sliced ok', e6.perft_profile()['launches'])\n")
    env = dict(os.environ, TAK_PERFT_CAP="20000")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "sliced ok" in out.stdout, out.stdout + out.stderr
    assert int(out.stdout.split()[-1]) >= 20         # an unsliced 6x6 perft(4) is 13 launches; the 132 720-state level alone is 7 slices
