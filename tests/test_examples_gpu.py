"""GPU parity of the batched device `Example::to_tensors` (alpha-tak/src/example.rs:63-78: 8-fold symmetry augmentation
of replay records into training tensors) against the oracle's restatement, bit for bit."""
import numpy as np
import pytest

import oracle
import tak_b200 as tb
from util import random_positions, splitmix, to_tb_state

pytestmark = pytest.mark.gpu


def _record(g, gi, result):
    rec = tb.ReplayRecord()
    rec.state = to_tb_state(g.state())
    rec.result = result
    moves = g.possible_moves()
    pol = [(m, 1 + splitmix(gi * 131 + i) % 97 if i % 3 else 0) for i, m in enumerate(moves)]
    rec.n_children = len(pol)
    for i, (m, v) in enumerate(pol):
        rec.moves[i], rec.visits[i] = m, v
    return rec, pol


@pytest.mark.parametrize("n,count", [(5, 40), (6, 64), (8, 6)])
def test_examples_to_tensors_match_oracle(n, count):
    games = random_positions(n, count, seed=17 + n, half_komi=4, max_ply=90)
    recs, pols = [], []
    for gi, g in enumerate(games):
        r, p = _record(g, gi, (1.0, 0.0, -1.0)[gi % 3])
        recs.append(r)
        pols.append(p)
    eng = tb.Engine(n, 8)
    inputs, pi, z = eng.examples_to_tensors(recs)
    assert inputs.shape == (8 * count, tb.input_channels(n), n, n) and pi.shape == (8 * count, tb.policy_size(n))
    for gi, g in enumerate(games):
        oi, op, oz = oracle.Example(g, pols[gi], recs[gi].result).to_tensors()
        sl = slice(8 * gi, 8 * gi + 8)
        assert np.array_equal(inputs[sl].view(np.uint32), oi.view(np.uint32)), f"inputs of example {gi}"
        assert np.array_equal(pi[sl].view(np.uint32), op.view(np.uint32)), f"pi of example {gi}"
        assert np.array_equal(z[sl], oz)
    # empty batch and a second, smaller call on the same engine (buffers are reused)
    a, b, c = eng.examples_to_tensors([])
    assert a.shape[0] == 0
    i2, p2, _ = eng.examples_to_tensors(recs[:3])
    assert np.array_equal(i2, inputs[:24]) and np.array_equal(p2, pi[:24])
    eng.close()


def test_selfplay_records_round_trip_through_text_and_tensors():
    """Records drained from the device self-play loop -> Example lines -> parsed back -> tensors: the replay path end to
    end (train/src/main.rs writes the lines, alpha-tak reads them back for training)."""
    n, G = 5, 64
    eng = tb.Engine(n, G, nodes_per_game=1 << 14, max_batch=G)
    eng.net_create(0)
    eng.selfplay_begin(rollouts=24, half_komi=4, instant_win=1, exploit_ply=8, noise_ply=0, seed=3, max_plies=60)
    recs = []
    for _ in range(70):
        eng.selfplay_step(1)
        recs += eng.selfplay_drain(16 * G)
        if len(recs) >= 200:
            break
    assert len(recs) >= 50, "no finished games"
    recs = recs[:200]
    lines = [tb.example_format(r) for r in recs]
    back = [tb.example_parse(l, n) for l in lines]
    for r, b, l in zip(recs, back, lines):
        assert tb.example_format(b) == l
        assert b.n_children == r.n_children and b.result == r.result and r.result in (1.0, 0.0, -1.0)
        # what the text cannot carry (example.rs:86: TPS drops reversible_plies) is all that may differ
        sa, sb = tb.TakState.from_buffer_copy(bytes(r.state)), tb.TakState.from_buffer_copy(bytes(b.state))
        sa.reversible_plies = 0
        assert bytes(sa) == bytes(sb)
    i1, p1, z1 = eng.examples_to_tensors(recs)
    i2, p2, z2 = eng.examples_to_tensors(back)
    assert np.array_equal(i1, i2) and np.array_equal(p1, p2) and np.array_equal(z1, z2)
    assert np.allclose(p1.sum(axis=1), 1.0, atol=1e-5)
    eng.close()
