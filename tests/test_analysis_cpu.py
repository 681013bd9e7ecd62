"""Host-side mirror of alpha_tak::{MoveInfo, NodeDebugInfo, Analysis} (debug.rs:42-106, analysis.rs) and the oracle's
Node::debug / continuation restatement -- CPU only (move text comes from the C ABI's host-side tak_ptn_format)."""
import oracle
import tak_b200 as tb
from tak_b200.analysis import Analysis, MoveInfo, NodeDebugInfo


def mv(text, n=6):
    return tb.parse_move(text, n)


def test_analysis_start_as_black_reference_vector():
    # alpha-tak/src/analysis.rs:265-288 (the reference's own test)
    a = Analysis(6, 4, 5)
    a.add_move(mv("Se4"), MoveInfo(mv("Se4"), 0, -1.0, 1.0, [], 6), 0.0)
    a.add_move_without_info(mv("c6"))
    a.add_move_without_info(mv("e4+"))
    assert str(a) == '[Size "6"]\n[Komi "2"]\n3. -- Se4 {r: +1.000, p: 1.0000, v: 0}\n4. c6 e4+\n'


def test_analysis_komi_and_settings():
    assert Analysis(5, 5, 0).settings == '[Size "5"]\n[Komi "2.5"]\n'
    assert Analysis(5, -3, 0).settings == '[Size "5"]\n[Komi "-1.5"]\n'      # Rust: -3 / 2 == -1, -3 % 2 == -1
    a = Analysis(6, 0, 0)
    a.add_setting("Player1", "new")
    assert a.settings == '[Size "6"]\n[Komi "0"]\n[Player1 "new"]\n'


def test_analysis_update_marks_branches_and_text():
    n = 6
    a = Analysis(n, 4, 0)
    d0 = NodeDebugInfo([MoveInfo(mv("a1"), 90, 0.5, 0.4, [(mv("f6"), 20000), (mv("b2"), 12000), (mv("c3"), 5)], n),
                        MoveInfo(mv("b1"), 85, 0.1, 0.3, [(mv("a6"), 11000)], n),
                        MoveInfo(mv("c1"), 10, -0.2, 0.3, [], n)])
    ev0 = d0.eval()
    a.update(d0, mv("a1"))
    assert a.branches and a.branches[0][0] == 0 and a.branches[0][1].mov == mv("b1")   # 85 > 0.9 * 90
    d1 = NodeDebugInfo([MoveInfo(mv("f6"), 100, 0.6, 0.9, [], n)])
    a.update(d1, mv("f6"))
    # eval_diff = -(0.6 + ev0) <= -0.4  => the move before (ply 0) was a blunder
    assert ev0 > 0 and a.marks == [(0, "blunder")]
    text = str(a)
    lines = text.split("\n")
    assert lines[2] == "1. a1??{evaluation: -0.600} {r: +0.500, p: 0.4000, v: 90} f6 {r: -0.600, p: 0.9000, v: 100}"
    assert "{0_b1}" in text and "1. b1  {r: +0.100, p: 0.3000, v: 85} a6" in text
    assert str(a.without_branches()).count("{0_b1}") == 0


def test_node_debug_info_table():
    n = 5
    d = NodeDebugInfo([MoveInfo(mv("a1", n), 3, 0.25, 0.5, [(mv("b2", n), 2), (mv("3c3>12", n), 1)], n),
                       MoveInfo(mv("Cb1", n), 1, -1.0, 0.5, [], n)])
    assert abs(d.eval() - (0.25 * 0.75 - 1.0 * 0.25)) < 1e-7
    text = d.format()
    assert text.splitlines()[0] == "evaluation: -0.0625"
    assert text.splitlines()[1] == "turn      visited   reward   policy | continuation"
    assert text.splitlines()[2] == "a1              3  +0.2500   0.5000 | b2 3c3>12"
    assert len(d.format(1).splitlines()) == 3
    assert NodeDebugInfo([]).format() == "Node has no children"
    assert [m.reward for m in d.maybe_flip(True).moves] == [-0.25, 1.0]


def test_oracle_debug_is_sorted_and_follows_the_most_visited_line():
    g = oracle.Game(5, 0)
    for m in ("a1", "e5", "b2", "c3"):
        g.play(m)
    s = oracle.Search(5)
    s.rollouts_dummy(g, 400)
    info = s.debug(4)
    mvs, vis, _, _, _ = s.children()
    assert sorted((v for _, v, _, _, _ in info), reverse=True) == [v for _, v, _, _, _ in info]
    assert sorted(m for m, *_ in info) == sorted(mvs.tolist())
    # the continuation of the top move == re-rooting on it and asking pick_move repeatedly
    top = info[0]
    assert top[0] == s.pick_move()
    s.play(top[0])
    for cm, cv in top[4]:
        assert cm == s.pick_move()
        kids = dict(zip(s.children()[0].tolist(), s.children()[1].tolist()))
        assert kids[cm] == cv
        s.play(cm)
