"""GPU parity of the `alpha_tak::Player` mirror (alpha-tak/src/player.rs:23-199): pipelined batches -- the NEXT batch
of leaves is selected before the PREVIOUS one is evaluated and backed up -- against the same schedule driven on the
oracle's Node, fed the engine's own network outputs.  Trees are compared bit for bit after every call."""
import numpy as np
import pytest

import oracle
import tak_b200 as tb
from tak_b200 import weights as W
from util import to_tb_state

pytestmark = pytest.mark.gpu


class OraclePlayer:
    """player.rs on the CPU oracle, with the deterministic interleaving: request (select) first, then consume."""

    def __init__(self, n, eng, batch, game):
        self.n, self.eng, self.batch = n, eng, batch
        self.search = oracle.Search(n)
        self.game = game
        self.outstanding = []
        self.examples = []
        self.request()

    def request(self):
        before = self.search.pending()
        for _ in range(self.batch):
            self.search.virtual_rollout(self.game)
        self.outstanding.append(self.search.pending() - before)

    def consume(self):
        for _ in range(self.outstanding.pop(0)):
            pol, val = self.eng.policy_eval([to_tb_state(self.search.pending_state(0))])
            self.search.devirtualize(pol[0], float(val[0]))

    def rollout(self):
        self.request()
        self.consume()

    def play_move(self, move):
        self.consume()
        mv, vis, _, _, _ = self.search.children()
        self.examples.append((bytes(self.game.state()), list(zip(mv.tolist(), vis.tolist()))))
        self.search.play(move)
        self.game.play(move)
        self.request()


def _oracle_debug_info(osearch, n, depth):
    from tak_b200.analysis import MoveInfo, NodeDebugInfo
    return NodeDebugInfo([MoveInfo(m, v, r, p, list(c), n) for m, v, r, p, c in osearch.debug(depth)])


def _same_debug(eng, gid, osearch, n, depth):
    """mcts_debug == Node::debug(depth) (search/debug.rs:9-40): order, stats bit for bit, continuations."""
    d, o = eng.debug(gid, depth), _oracle_debug_info(osearch, n, depth)
    assert len(d.moves) == len(o.moves)
    for a, b in zip(d.moves, o.moves):
        assert (a.mov, a.visits, a.continuation) == (b.mov, b.visits, b.continuation)
        assert np.float32(a.reward).view(np.uint32) == np.float32(b.reward).view(np.uint32)
        assert np.float32(a.policy).view(np.uint32) == np.float32(b.policy).view(np.uint32)
    assert d.format(5) == o.format(5)
    return o


def _same_tree(eng, gid, osearch):
    mv, vis, pri, rew = eng.children(gid)
    omv, ovis, opri, orew, _ = osearch.children()
    assert np.array_equal(mv, omv) and np.array_equal(vis, ovis)
    assert np.array_equal(pri.view(np.uint32), opri.view(np.uint32))
    assert np.array_equal(rew.view(np.uint32), orew.view(np.uint32))
    assert eng.root(gid)[0] == osearch.root()[0]


@pytest.mark.parametrize("n,arch,batch", [(5, 5, 4), (6, 6, 8), (6, 0, 16)])
def test_player_pipelined_batches_match_oracle(n, arch, batch):
    eng = tb.Engine(n, 4, nodes_per_game=1 << 15, max_batch=64)
    eng.net_create(arch)
    if arch:
        eng.net_load_weights(W.random_weights(arch, seed=4))
    gid = 2                                   # other slots hold unrelated games with their own queued leaves
    eng.reset(0, 4, 4)
    eng.reserve_pending(2 * batch)
    eng.tree_reset([0, 1, 3])
    eng.virtual_rollout([0, 3], 1)            # must survive every partial devirtualize below
    g = oracle.Game(n, 4)
    for m in ("a1", f"{'abcdefgh'[n - 1]}{n}"):
        g.play(m)
    from tak_b200.analysis import Analysis, MAX_BRANCH_LENGTH
    p = tb.Player(eng, gid, batch, save_examples=True, state=to_tb_state(g.state()), create_analysis=True)
    o = OraclePlayer(n, eng, batch, g.clone())
    o_analysis = Analysis(n, 4, 2)
    plies = 0
    while o.game.result() == 0 and plies < 6:
        for _ in range(5):
            p.rollout()
            o.rollout()
            # trees carry the virtual visits of the still-outstanding batch on both sides
            _same_tree(eng, gid, o.search)
        for depth in (0, 1, 3, MAX_BRANCH_LENGTH, 16):
            _same_debug(eng, gid, o.search, n, depth)
        mv = p.pick_move(True)
        assert mv == o.search.pick_move()
        p.play_move(mv)
        o.consume()                               # play_move's "rollout stale paths" happens before the analysis update
        o.outstanding.insert(0, 0)
        o_analysis.update(_oracle_debug_info(o.search, n, MAX_BRANCH_LENGTH), mv)
        o.play_move(mv)
        _same_tree(eng, gid, o.search)
        assert eng.download([gid])[0].key() == bytes(o.game.state())
        plies += 1
    gids, _ = eng.pending(with_states=False)
    assert (gids == 0).sum() == 1 and (gids == 3).sum() == 1, "another game's queued leaf was consumed"
    assert len(p.examples) == len(o.examples) == plies
    for (st, pol), (ost, opol) in zip(p.examples, o.examples):
        assert bytes(st) == ost and pol == opol
    text = str(p.get_analysis())
    assert text == str(o_analysis) and text.startswith(f'[Size "{n}"]\n[Komi "2"]\n2. ')
    assert p.analysis.played_moves == []
    recs = p.get_examples(tb.RESULT_WHITE | tb.RESULT_FLAG)
    assert [r.result for r in recs] == [1.0 if r.state.to_move == 0 else -1.0 for r in recs]
    assert p.examples == []
    eng.close()
