"""Statistical GPU tests of the two reference paths that consume `thread_rng` and therefore cannot be pinned bit for
bit: Node::apply_dirichlet (alpha-tak/src/search/noise.rs:6-16) and Node::pick_move(false) (search/play.rs:60-65).
The device draws from a counter-based generator (splitmix64 keyed by seed x game x child); what IS checkable is checked:
the noise is a Dirichlet(alpha) sample per root (sums to 1, mean 1/k, component variance), the 0.7/0.3 mix is applied to
the existing priors, the "no visit yet" panic surfaces as an error, and sampled picks follow visits / sum(visits)."""
import numpy as np
import pytest

import tak_b200 as tb
from tak_b200 import weights as W

pytestmark = pytest.mark.gpu

ALPHA, RATIO = 0.2, 0.3          # NOISE_ALPHA / NOISE_RATIO, train/src/self_play.rs:14-15


def _same_position_engine(n, G, arch, plies=("a1", "e5", "c3", "c2")):
    eng = tb.Engine(n, G, nodes_per_game=1 << 15, max_batch=G)
    eng.net_create(arch)
    if arch:
        eng.net_load_weights(W.random_weights(arch, seed=4))
    ids = np.arange(G, dtype=np.int32)
    eng.reset(0, G, 4)
    for m in plies:
        assert not eng.play(ids, [tb.parse_move(m, n)] * G).any()
    eng.tree_reset(ids)
    return eng, ids


def test_dirichlet_noise_is_a_dirichlet_sample_mixed_into_the_priors():
    n, G = 5, 1536
    eng, ids = _same_position_engine(n, G, 0)
    # noise.rs:7-10: "cannot apply dirichlet noise without initialized policy"
    with pytest.raises(tb.TakNativeError) as ei:
        eng.apply_dirichlet(ids[:4], ALPHA, RATIO, 1)
    assert ei.value.code == -37
    eng.rollouts(ids, 1)                              # Node::rollout once (self_play.rs:177): DummyNet priors are all 1.0
    mv0, _, pri0, _ = eng.children(0)
    k = len(mv0)
    assert k > 40 and np.all(pri0 == 1.0)
    eng.apply_dirichlet(ids, ALPHA, RATIO, 0xD1CE)
    eta = np.zeros((G, k), dtype=np.float64)
    for gid in range(G):
        mv, _, pri, _ = eng.children(gid)
        assert np.array_equal(mv, mv0)
        eta[gid] = (pri.astype(np.float64) - (1.0 - RATIO) * 1.0) / RATIO     # prior' = eta*ratio + prior*(1-ratio)
    assert np.all(eta > -1e-6)
    assert np.allclose(eta.sum(axis=1), 1.0, atol=2e-5), np.abs(eta.sum(axis=1) - 1).max()
    # Dirichlet(alpha,...,alpha): E = 1/k, Var = (1/k)(1-1/k)/(k*alpha+1)
    var = (1.0 / k) * (1.0 - 1.0 / k) / (k * ALPHA + 1.0)
    se = np.sqrt(var / G)
    assert np.abs(eta.mean(axis=0) - 1.0 / k).max() < 5.5 * se, (np.abs(eta.mean(axis=0) - 1.0 / k).max(), se)
    emp = eta.var(axis=0).mean()
    assert abs(emp / var - 1.0) < 0.1, (emp, var)
    # alpha = 0.2 is a spiky prior: most of the mass of a sample sits on a few children
    assert np.median(eta.max(axis=1)) > 4.0 / k
    # independent across games and children, reproducible for a seed, different for another seed
    assert len({eta[g].tobytes() for g in range(G)}) == G
    c = np.corrcoef(eta[:, 0], eta[:, 1])[0, 1]
    assert abs(c + 1.0 / (k - 1)) < 0.12               # Dirichlet components are weakly anti-correlated: -1/(k-1)
    eng2, _ = _same_position_engine(n, 8, 0)
    eng2.rollouts(ids[:8], 1)
    eng2.apply_dirichlet(ids[:8], ALPHA, RATIO, 0xD1CE)
    for gid in range(8):
        assert np.array_equal(eng2.children(gid)[2], eng.children(gid)[2])
    eng2.apply_dirichlet(ids[:8], ALPHA, RATIO, 0xD1CF)
    assert not np.array_equal(eng2.children(0)[2], eng.children(0)[2])
    eng.close()
    eng2.close()


def test_dirichlet_mix_on_network_priors():
    """prior <- noise*ratio + prior*(1-ratio) on top of Net6 priors (not uniform, and summing to less than 1 because the
    softmax spans illegal moves too): (prior' - 0.7*prior)/0.3 must again be a point of the simplex."""
    n, G = 6, 64
    eng, ids = _same_position_engine(n, G, 6, plies=("a1", "f6", "c3", "d4"))
    eng.rollouts(ids, 1)
    before = [eng.children(g)[2].astype(np.float64) for g in range(G)]
    assert before[0].sum() < 1.0 and before[0].std() > 0
    eng.apply_dirichlet(ids, ALPHA, RATIO, 99)
    for g in range(G):
        after = eng.children(g)[2].astype(np.float64)
        eta = (after - (1.0 - RATIO) * before[g]) / RATIO
        assert np.all(eta > -1e-6) and abs(eta.sum() - 1.0) < 5e-5
    eng.close()


def test_sampled_pick_follows_visit_counts():
    """pick_move(false) = WeightedIndex over the children's visit counts (play.rs:60-65): chi-square of 16 384 draws (one per
    game; all games hold the same 300-rollout tree) against visits / sum(visits); unvisited children are never drawn."""
    n, G, R = 5, 16384, 300
    eng, ids = _same_position_engine(n, G, 0)
    eng.rollouts(ids, R)
    mv, vis, _, _ = eng.children(0)
    for gid in (1, G // 2, G - 1):
        assert np.array_equal(eng.children(gid)[1], vis)
    assert vis.sum() == R - 1
    p = vis / vis.sum()
    index = {int(m): i for i, m in enumerate(mv)}
    for seed in (5, 6):
        picks = eng.pick_move_sampled(ids, seed)
        obs = np.bincount([index[int(m)] for m in picks], minlength=len(mv))
        assert obs[vis == 0].sum() == 0
        live = vis > 0
        exp = p[live] * G
        chi2 = ((obs[live] - exp) ** 2 / exp).sum()
        dof = live.sum() - 1
        assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)
        # not the argmax in disguise: the most visited child is drawn about p_max of the time
        assert abs(obs.max() / G - p.max()) < 0.02
    a, b = eng.pick_move_sampled(ids, 5), eng.pick_move_sampled(ids, 6)
    assert np.array_equal(a, eng.pick_move_sampled(ids, 5)) and not np.array_equal(a, b)
    eng.close()
