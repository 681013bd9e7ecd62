"""GPU parity of the network path: game_repr planes (bit-exact vs the oracle) and Net5/Net6 policy_eval
(within the north-star tolerance, max abs 1e-2, vs the fp32 PyTorch restatement on seeded random weights)."""
import numpy as np
import pytest
import torch

import oracle
import tak_b200 as tb
from oracle.net_ref import RefNet
from tak_b200 import weights as W
from util import random_positions, to_tb_state

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-2    # north_star: "network policy logits and value agree within a stated bf16 tolerance (max abs 1e-2)"
VALUE_TOL = 1e-2
POLICY_REL_TOL = 1e-2   # softmax probabilities: max |p - p_ref| / max p_ref (an absolute bound is vacuous at p ~ 1e-4)


@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 8])
def test_game_repr_bit_exact(n):
    # reference: alpha-tak/src/repr/game.rs:19-51 (+ tests.rs:20-111 through the oracle, pinned on CPU)
    games = random_positions(n, 24, seed=n, max_ply=90)
    eng = tb.Engine(n, 8)
    got = eng.game_repr([to_tb_state(g.state()) for g in games])
    want = np.stack([g.repr() for g in games])
    assert got.shape == want.shape == (24, tb.input_channels(n), n, n)
    assert np.array_equal(got, want)
    eng.close()


def _check_net(arch, batch):
    n = arch
    blob = W.random_weights(arch, seed=0)
    games = random_positions(n, batch, seed=11 + arch)
    states = [to_tb_state(g.state()) for g in games]
    eng = tb.Engine(n, 8, max_batch=4096)
    eng.net_create(arch)
    assert eng.net_weights_size() == blob.size == W.blob_size(arch)
    eng.net_load_weights(blob)
    pol, val = eng.policy_eval(states)
    lg, val_l = eng.policy_logits(states)
    ref = RefNet(arch, blob, device="cuda")          # fp32: RefNet switches TF32 off for cuDNN and cuBLAS
    assert not torch.backends.cudnn.allow_tf32 and not torch.backends.cuda.matmul.allow_tf32
    x = torch.from_numpy(np.stack([g.repr() for g in games])).cuda()
    rpol, rval, rlogits = ref.forward_mcts(x)
    rpol, rval, rlogits = rpol.cpu().numpy(), rval.cpu().numpy(), rlogits.cpu().numpy()
    lerr = np.abs(lg - rlogits).max()
    verr = np.abs(val - rval).max()
    rel = np.abs(pol - rpol).max() / rpol.max()
    print(f"Net{arch} B={batch}: logits max|err| {lerr:.3e} (|logit| max {np.abs(rlogits).max():.3f}), policy rel "
          f"{rel:.3e} (max p {rpol.max():.3e}), value max|err| {verr:.3e}, |value| max {np.abs(rval).max():.3f}")
    assert np.array_equal(val, val_l)
    assert np.allclose(pol.sum(1), 1.0, atol=1e-3)
    # the tolerance is on the LOGITS, as north_star words it
    assert lerr < LOGIT_TOL and verr < VALUE_TOL
    assert rel < POLICY_REL_TOL
    # policy_eval is softmax(policy_logits) over the whole vector
    sm = np.exp(lg - lg.max(axis=1, keepdims=True))
    sm /= sm.sum(axis=1, keepdims=True)
    assert np.abs(sm - pol).max() / pol.max() < 1e-5
    # batch independence: the same position evaluates to the same bits wherever it sits in a batch
    pol2, val2 = eng.policy_eval(states[::-1][: max(1, batch // 3)])
    k = pol2.shape[0]
    assert np.array_equal(pol2, pol[::-1][:k]) and np.array_equal(val2, val[::-1][:k])
    eng.close()


def test_net6_policy_eval():
    _check_net(6, 96)


def test_net5_policy_eval():
    _check_net(5, 64)


def test_net6_large_batch_matches_small():
    blob = W.random_weights(6, seed=3)
    games = random_positions(6, 700, seed=5)
    states = [to_tb_state(g.state()) for g in games]
    eng = tb.Engine(6, 8, max_batch=1024)
    eng.net_create(6)
    eng.net_load_weights(blob)
    pol, val = eng.policy_eval(states)           # one 700-board pass (many tiles per CTA)
    pol_s, val_s = eng.policy_eval(states[:5])   # 5-board pass
    assert np.array_equal(pol[:5], pol_s) and np.array_equal(val[:5], val_s)
    eng.close()


def test_dummy_net():
    # reference: alpha-tak/src/search/tests.rs:29-34
    eng = tb.Engine(3, 4)
    eng.net_create(0)
    pol, val = eng.policy_eval([tb.state_init(3)])
    assert pol.shape == (1, tb.policy_size(3)) and np.all(pol == 1.0) and val[0] == 0.0
    eng.close()


def test_board_repr_reference_plane_dump(golden):
    """alpha-tak/src/repr/tests.rs:20-111 through the DEVICE kernel: the reference dumps board_repr(game, White) of a
    45-ply 5x5 game; game_repr encodes from the side to move (Black at ply 45), i.e. the same planes with every
    mine / theirs channel pair swapped."""
    r = golden["board_repr"]
    game = tb.Game.from_ptn_moves(r["n"], r["moves"])
    st = game.state()
    assert st.ply == 45 and st.to_move == 1 and r["to_move_arg"] == 0
    planes = game.engine.game_repr([st])[0]
    want = np.array(r["planes_12x5x5"], dtype=np.float32).reshape(12, 5, 5)
    for ch in range(12):
        assert np.array_equal(planes[ch ^ 1], want[ch]), ch
    # the deeper stack planes of this position are empty in the reference dump (tests.rs:107-110)
    assert not planes[12:2 * (5 + 8)].any()
    # empty board: all board planes zero (tests.rs:11-17)
    empty = tb.Game.default(5)
    assert not empty.engine.game_repr([empty.state()])[0][:2 * (5 + 8)].any()
