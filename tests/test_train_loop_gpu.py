"""The `train` binary's loop (train/src/main.rs:82-123) end to end on the device at toy sizes: self-play produces replay
records, Network::train consumes them (shuffled, chunks_exact, a step every CHUNKS_IN_STEP chunks), the candidate is
pitted against the current network, self-play continues with whichever network was kept."""
import numpy as np
import pytest

import tak_b200 as tb
from tak_b200 import train_loop as TL
from tak_b200 import weights as W

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [6, 5])
def test_training_loop_iteration(tmp_path, n):
    G = 64
    blob = W.random_weights(n, seed=5)
    cur = tb.Engine(n, G, nodes_per_game=1 << 12, max_batch=G)
    cand = tb.Engine(n, G, nodes_per_game=1 << 12, max_batch=G)
    for e in (cur, cand):
        e.net_create(n)
        e.net_load_weights(blob)
    sp = dict(rollouts=12, half_komi=4, instant_win=1, exploit_ply=6, noise_ply=8, max_plies=24, seed=3)
    lines = []
    rng = np.random.default_rng(0)
    # first turn: no examples yet -> self-play only
    blob1, examples, res = TL.training_iteration(cur, cand, blob, [], rng, min_new_examples=200, selfplay_kw=sp,
                                                 log=lines.append, save_dir=str(tmp_path))
    import os
    data = [f for f in os.listdir(tmp_path / "_examples") if f.endswith(".data")]
    assert len(data) == 1
    with open(tmp_path / "_examples" / data[0]) as f:
        text = f.read().split("\n")
    assert len(text) - 1 == len(examples) and tb.example_parse(text[0], n).n_children == examples[0].n_children
    assert res is None and np.array_equal(blob1, blob) and len(examples) >= 200
    assert all(r.n_children > 0 and r.result in (-1.0, 0.0, 1.0) for r in examples)
    # the replay text format round-trips every record (example.rs:81-133)
    for r in examples[:20]:
        back = tb.example_parse(tb.example_format(r), n)
        # the text carries the position as TPS: komi and the reversible-ply counter are not part of it (tps.rs:37-96)
        assert tb.tps_format(back.state) == tb.tps_format(r.state)
        assert back.n_children == r.n_children and back.result == r.result
        assert list(back.moves[:r.n_children]) == list(r.moves[:r.n_children])
        assert list(back.visits[:r.n_children]) == list(r.visits[:r.n_children])
    # second turn: train (chunks of 16 examples x 8 symmetries, a step every 3 chunks), pit, self-play
    n_chunks = len(examples) // 16
    blob2, examples2, res = TL.training_iteration(
        cur, cand, blob1, examples, rng, pit_games=2, pit_rollouts=2, pit_batch=4, min_new_examples=50,
        train_kw=dict(chunk_size=16, chunks_in_step=3, lr=1e-3), selfplay_kw=sp, log=lines.append)
    assert lines.count("making step!") == n_chunks // 3
    assert sum(1 for s in lines if s.startswith("p=")) == n_chunks
    assert res is not None and res.wins + res.losses + res.draws == 4
    losses = TL.train_network.last_losses
    assert all(np.isfinite(l).all() for l in losses)
    first, last = np.mean([sum(l) for l in losses[:3]]), np.mean([sum(l) for l in losses[-3:]])
    print("\ntraining loss, first / last 3 chunks:", first, last)
    assert last < first, (first, last)                    # the steps taken reduce the training loss
    accepted = res.win_rate() > TL.WIN_RATE_THRESHOLD
    assert np.array_equal(blob2, blob1) != accepted
    assert len(examples2) >= len(examples) + 50
    cur.close()
    cand.close()
