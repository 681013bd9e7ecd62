"""CPU checks of oracle/selfplay_ref.py, the restatement of train::self_play_parallel (train/src/self_play.rs:96-262)
that the GPU self-play loop is compared with: the instant-win scan, its 1000/1 fake visits, example completion from the
mover's perspective, slot restart, and the reference's restart-at-ply-0 quirk."""
import oracle
from oracle.selfplay_ref import SelfPlayParallel, result_to_number


def test_instant_win_scan_marks_exactly_the_winning_moves():
    # tak/tests/wins.rs-style road position: white to move completes a1-a2-a3 with a3
    g = oracle.Game.from_ptn_moves(3, ["a3", "c1", "a1", "c2", "a2", "b2"], half_komi=0)
    assert g.result() == 0
    policy, win = g.instant_win_policy()
    assert [m for m, _ in policy] == g.possible_moves()
    for mv, visits in policy:
        c = g.clone()
        c.play(mv)
        r = c.result()
        wins = (r & 0xF) == (1 if g.state().to_move == 0 else 2)
        assert visits == (1000 if wins else 1)
    assert win == any(v == 1000 for _, v in policy)
    assert not oracle.Game(3, 0).instant_win_policy()[1]


def test_loop_completes_examples_and_keeps_the_restart_quirk():
    n, G, R = 3, 8, 30
    sp = SelfPlayParallel(n, G, R, policy_eval=None, coin=lambda slot, serial: (slot + serial) % 2 == 0, exploit_plies=0)
    quirk = 0
    for _ in range(30):
        before = sp.completed_games
        serial = list(sp.serial)
        sp.iteration()
        for i in range(G):
            ply = sp.games[i].state().ply
            assert ply != 0 or sp.serial[i] == serial[i] + 1      # ply 0 only right after a normal game end
        if sp.completed_games > before:
            quirk += sum(1 for i in range(G) if sp.games[i].state().ply == 1)
    assert sp.completed_games >= G
    assert quirk > 0, "no slot was searched from the empty board after an instant win"
    instant = 0
    for e in sp.examples:
        g = oracle.Game.from_state(e.state)
        assert [m for m, _ in e.policy] == g.possible_moves()
        assert e.result in (1.0, 0.0, -1.0)
        if any(v == 1000 for _, v in e.policy):
            instant += 1
            assert all(v in (1, 1000) for _, v in e.policy) and e.result == 1.0   # the mover wins on the spot
    assert instant > 0
    # the examples of one finished game alternate perspective with the side to move
    by_game = {}
    for e in sp.examples:
        by_game.setdefault((e.slot, e.serial), []).append(e)
    for exs in by_game.values():
        white = {e.result if e.state.to_move == 0 else -e.result for e in exs}
        assert len(white) == 1
    assert result_to_number(0x11) == 1.0 and result_to_number(2) == -1.0 and result_to_number(0x13) == 0.0
