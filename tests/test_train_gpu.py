"""GPU parity of the training step (Network::train_inner + Adam, alpha-tak/src/model/network.rs:37-97,
forward_training net6.rs:111-122) against the fp32 PyTorch restatement oracle/net_ref.py:RefTrainer on the same weights
and the same chunk of examples.  The device path computes in bf16 (operands and saved activations) with fp32
accumulation, fp32 master weights and fp32 gradients; tolerances are written next to each check."""
import numpy as np
import pytest

import oracle
import tak_b200 as tb
from oracle.net_ref import RefTrainer
from tak_b200 import weights as W
from util import random_positions, splitmix, to_tb_state

pytestmark = pytest.mark.gpu


def make_chunk(n_pos, seed, n=6):
    """(inputs [b,C,n,n], pi [b,policy_size], z [b]) from oracle positions: pi = normalised pseudo-random visit counts over
    the legal moves, z in {-1, 0, 1}."""
    games = random_positions(n, n_pos, seed=seed, max_ply=60)
    P = oracle.policy_size(n)
    x = np.stack([g.repr() for g in games]).astype(np.float32)
    pi = np.zeros((n_pos, P), dtype=np.float32)
    z = np.zeros(n_pos, dtype=np.float32)
    for i, g in enumerate(games):
        moves = g.possible_moves()
        v = np.array([1 + splitmix(seed * 31 + i * 1009 + k) % 50 for k in range(len(moves))], dtype=np.float64)
        v[splitmix(seed + i) % len(moves)] += 400                         # a peaked visit distribution
        for m, c in zip(moves, v / v.sum()):
            pi[i, oracle.move_index(m, n)] = c
        z[i] = float(splitmix(seed * 17 + i) % 3) - 1.0
    return x, pi, z


def rel_err(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(1e-30, np.linalg.norm(b.astype(np.float64))))


@pytest.fixture(scope="module", params=[6, 5], ids=["net6", "net5"])
def trained(request):
    arch = request.param
    blob = W.random_weights(arch, seed=11)
    eng = tb.Engine(arch, 8, nodes_per_game=1 << 10, max_batch=8)
    eng.net_create(arch)
    eng.net_load_weights(blob)
    eng.train_begin(256)
    ref = RefTrainer(arch, blob, device="cuda")
    emu = RefTrainer(arch, blob, device="cuda", emulate_bf16=True)
    chunks = [make_chunk(200, 3, arch), make_chunk(131, 4, arch)]         # ragged: the second chunk is smaller
    losses, ref_losses = [], []
    for x, pi, z in chunks:
        losses.append(eng.train_chunk(x, pi, z))
        ref_losses.append(ref.chunk(x, pi, z))
        emu.chunk(x, pi, z)
    out = {"arch": arch, "eng": eng, "ref": ref, "blob": blob, "losses": losses, "ref_losses": ref_losses,
           "grads": eng.train_get(1), "ref_grads": ref.grads(), "emu_grads": emu.grads(), "chunks": chunks}
    yield out
    eng.close()


def test_losses_match(trained):
    # bf16 forward through 33 convolutions: |loss - ref| <= 2e-2 * |ref| (measured ~3e-3)
    for (lp, lz), (rp, rz) in zip(trained["losses"], trained["ref_losses"]):
        assert abs(lp - rp) <= 2e-2 * abs(rp), (lp, rp)
        assert abs(lz - rz) <= 2e-2 * max(abs(rz), 0.1), (lz, rz)


def test_gradients_match_autograd(trained):
    """Accumulated gradient (2 chunks) per tensor, relative L2 error against autograd.

    Measured (this test prints the table): the heads agree with fp32 autograd to 1e-3 .. 8e-3; going down the tower the
    error grows by about 1.5 % per residual block to ~0.27 at the first layers -- and a PyTorch autograd run that merely
    ROUNDS the stored activations / activation gradients / conv operands to bf16 (RefTrainer(emulate_bf16=True)) is
    just as far from fp32 (0.27): ReLU masks and BatchNorm statistics of a 33-layer random-init tower amplify bf16
    storage noise.  So the criterion is: the device path is as close to fp32 autograd as bf16-rounding autograd is
    (<= 1.5x its error + 3e-2), the heads are tight, and the gradient direction is kept (cosine >= 0.93 per tensor)."""
    arch = trained["arch"]
    g = W.split(trained["grads"], arch)
    r = W.split(trained["ref_grads"], arch)
    m = W.split(trained["emu_grads"], arch)
    print("\nrel L2 error per tensor: device vs fp32 | device vs bf16-emulating autograd | emulation vs fp32 | cosine")
    bad = {}
    for name in g:
        if name.endswith("running_mean") or name.endswith("running_var"):
            continue
        if name.endswith("conv1.bias") or name.endswith("conv2.bias") or name == "initial_conv.bias":
            # a conv bias feeding a batch-statistics BatchNorm has a mathematically zero gradient: autograd returns fp32
            # rounding noise, the device path exactly 0
            assert float(np.abs(g[name]).max()) == 0.0
            assert float(np.abs(r[name]).max()) <= 1e-4 * float(np.abs(r["value_fc.weight"]).max() + 1.0)
            continue
        e_dev, e_emu, e_de = rel_err(g[name], r[name]), rel_err(m[name], r[name]), rel_err(g[name], m[name])
        a, b = g[name].astype(np.float64).ravel(), r[name].astype(np.float64).ravel()
        cos = float(a @ b / max(1e-300, np.linalg.norm(a) * np.linalg.norm(b)))
        print(f"  {name:28s} {e_dev:9.4f} {e_de:9.4f} {e_emu:9.4f} {cos:9.4f}")
        if not (e_dev <= 1.5 * e_emu + 3e-2 and e_de <= 1.5 * e_emu + 3e-2 and cos >= 0.93):
            bad[name] = (e_dev, e_de, e_emu, cos)
    assert not bad, f"gradient mismatch: {bad}"
    head = "policy_conv" if arch == 6 else "policy_fc"
    for name in (head + ".weight", head + ".bias", "value_fc.weight", "value_fc.bias"):
        assert rel_err(g[name], r[name]) <= 2e-2, name
    # the last residual block sits one step behind the heads: still tight
    last = f"block{15 if arch == 6 else 7}"
    for name in (last + ".conv2.weight", last + ".bn2.weight", last + ".bn2.bias"):
        assert rel_err(g[name], r[name]) <= 5e-2, name


@pytest.mark.parametrize("arch", [6, 5])
def test_training_reduces_the_loss_like_the_reference(arch):
    """End to end: 12 Adam steps (lr 1e-3) on one fixed chunk drive the loss down on the device path as they do in the
    fp32 reference (same start, same data): both fall by > 25 %, the first three losses agree to 1.5 %, and the last one
    is within [-20 %, +10 %] of the reference's.  The wide last bound is what the experiment allows, not slack for the
    kernels: at ten times the reference's learning rate on 160 positions the trajectory is chaotic -- the fp32 reference
    (cuDNN atomics) ends anywhere in 5.18 .. 5.39 on Net5 from run to run, and bf16 storage noise moves the device path
    further (measured over seven runs and three builds of the backward pass: 0.89 .. 0.97 of the reference, always on
    the low side); Net6 follows the reference to 0.4 % on every step."""
    blob = W.random_weights(arch, seed=21)
    x, pi, z = make_chunk(160, 9, arch)
    eng = tb.Engine(arch, 8, nodes_per_game=1 << 10, max_batch=8)
    eng.net_create(arch)
    eng.net_load_weights(blob)
    eng.train_begin(160)
    ref = RefTrainer(arch, blob, device="cuda", lr=1e-3, wd=1e-4)
    dev_loss, ref_loss = [], []
    for _ in range(12):
        dev_loss.append(sum(eng.train_chunk(x, pi, z)))
        eng.train_step(1e-3, 1e-4)
        ref_loss.append(sum(ref.chunk(x, pi, z)))
        ref.step()
    print("\nloss device:", [round(v, 3) for v in dev_loss], "\nloss fp32  :", [round(v, 3) for v in ref_loss])
    assert dev_loss[-1] < 0.75 * dev_loss[0] and ref_loss[-1] < 0.75 * ref_loss[0]
    for i in range(3):
        assert abs(dev_loss[i] - ref_loss[i]) <= 0.015 * ref_loss[i], (i, dev_loss[i], ref_loss[i])
    assert 0.80 * ref_loss[-1] <= dev_loss[-1] <= 1.10 * ref_loss[-1]
    eng.train_end()
    eng.close()


def test_running_statistics_updated(trained):
    w = W.split(trained["eng"].train_get(0), trained["arch"])
    r = W.split(trained["ref"].blob(), trained["arch"])
    for name in w:
        if name.endswith("running_mean"):
            assert np.abs(w[name] - r[name]).max() <= 2e-2 * max(1.0, float(np.abs(r[name]).max())), name
        if name.endswith("running_var"):
            assert rel_err(w[name], r[name]) <= 3e-2, name


def test_adam_step_matches_torch_on_the_same_gradients(trained):
    """Adam itself is exact arithmetic on fp32: feed torch.optim.Adam the DEVICE gradients and compare the updated weights
    (two steps, so the moment estimates and bias corrections are exercised): max abs diff <= 2e-7."""
    eng, blob, arch = trained["eng"], trained["blob"], trained["arch"]
    ref = RefTrainer(arch, blob, device="cuda")
    for step in range(2):
        if step == 1:
            x, pi, z = trained["chunks"][1]
            eng.train_chunk(x, pi, z)
        ref.set_grads(eng.train_get(1))
        # keep the reference's BN running statistics out of the comparison: only trainable tensors are stepped
        eng.train_step(1e-4, 1e-4)
        ref.step()
        assert float(np.abs(eng.train_get(1)).max()) == 0.0               # zero_grad
        w, r = W.split(eng.train_get(0), arch), W.split(ref.blob(), arch)
        for name in w:
            if "running_" in name:
                continue
            assert float(np.abs(w[name] - r[name]).max()) <= 2e-7, (step, name)
    st = eng.train_stats()
    assert st["steps"] == 2 and st["chunks_pending"] == 0
    # the stepped weights can be searched with: load them into the inference path
    new_blob = eng.train_get(0)
    assert np.isfinite(new_blob).all() and float(np.abs(new_blob - blob).max()) > 0
    eng.net_load_weights(new_blob)
    g = oracle.Game(arch, 4)
    pol, val = eng.policy_eval([to_tb_state(g.state())])
    assert abs(float(pol.sum()) - 1.0) < 1e-3 and np.isfinite(val).all()


def test_device_tensors_and_grad_view(trained):
    """net_train_chunk with device pointers == with host pointers; the zero-copy gradient tensor aliases the blob."""
    import torch
    eng = trained["eng"]
    x, pi, z = trained["chunks"][0]
    a = eng.train_chunk(x, pi, z)
    g_host = eng.train_get(1)
    gt = eng.train_grad_tensor()
    assert gt.shape[0] == g_host.size and np.array_equal(gt.cpu().numpy(), g_host)
    gt.zero_()
    assert float(np.abs(eng.train_get(1)).max()) == 0.0
    b = eng.train_chunk(torch.from_numpy(x).cuda(), torch.from_numpy(pi).cuda(), torch.from_numpy(z).cuda())
    assert a[0] == b[0] and a[1] == b[1]
    # not bit-identical run to run: the split-K partial sums and double atomics are order dependent
    assert rel_err(eng.train_get(1), g_host) < 1e-3
    gt.zero_()


@pytest.mark.parametrize("arch", [6, 5])
def test_two_pass_batchnorm_backward_agrees_with_the_fused_sums(arch, tmp_path):
    """The dgrad epilogue's BatchNorm-backward sums (conv_tc3.cuh ConvParams::bnb_y, the default) against the round-1 route
    that stays in the library as a fallback (TAK_TRAIN_BNB=0: k_bn_bwd_reduce + masking apply pass): same weights, same
    chunk.  Both paths store the same bf16 masked gradient and differ in the summation order of the per-channel sums only,
    so the last block's conv gradients agree to 1e-4 (measured 2e-7 / 3e-6 for Net6 / Net5); a last-bit difference in a
    sum flips a few bf16 roundings of dy, and 33 layers of masks and batch statistics amplify those (as they amplify
    bf16 against fp32), so every tensor is held to 3 % (measured worst: 1.2 % at Net6's block0.bn1.bias, 0.7 % at
    Net5's first conv).  The switch is read once per process, so the fallback runs in a child process."""
    import os
    import subprocess
    import sys
    import textwrap
    blob = W.random_weights(arch, seed=5)
    x, pi, z = make_chunk(150, 13, arch)
    np.savez(tmp_path / "chunk.npz", blob=blob, x=x, pi=pi, z=z)
    eng = tb.Engine(arch, 8, nodes_per_game=1 << 10, max_batch=8)
    eng.net_create(arch)
    eng.net_load_weights(blob)
    eng.train_begin(150)
    loss = eng.train_chunk(x, pi, z)
    grads = eng.train_get(1)
    eng.close()
    child = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
        import tak_b200 as tb
        d = np.load({str(tmp_path / "chunk.npz")!r})
        eng = tb.Engine({arch}, 8, nodes_per_game=1 << 10, max_batch=8)
        eng.net_create({arch}); eng.net_load_weights(d["blob"]); eng.train_begin(150)
        loss = eng.train_chunk(d["x"], d["pi"], d["z"])
        np.savez({str(tmp_path / "out.npz")!r}, grads=eng.train_get(1), loss=np.array(loss))
        eng.close()
    """)
    env = dict(os.environ, TAK_TRAIN_BNB="0")
    subprocess.run([sys.executable, "-c", child], check=True, env=env, timeout=300)
    out = np.load(tmp_path / "out.npz")
    assert abs(out["loss"][0] - loss[0]) <= 1e-5 * abs(loss[0]) and abs(out["loss"][1] - loss[1]) <= 1e-5 * max(abs(loss[1]), 0.1)
    g2 = W.split(out["grads"], arch)
    worst = ("", 0.0)
    for name, g in W.split(grads, arch).items():
        if np.abs(g).max() == 0.0 and np.abs(g2[name]).max() == 0.0:
            continue                                      # conv biases under a batch-statistics BatchNorm, running stats
        worst = max(worst, (name, rel_err(g, g2[name])), key=lambda t: t[1])
    last = f"block{15 if arch == 6 else 7}"
    last_err = max(rel_err(g, g2[name]) for name, g in W.split(grads, arch).items() if name.startswith(last + ".conv"))
    print("\nworst tensor:", worst, "last block convs:", last_err)
    assert worst[1] <= 3e-2, worst
    assert last_err <= 1e-4, last_err


def test_training_api_errors():
    """Error behaviour at the boundary: status codes, never aborts."""
    e4 = tb.Engine(4, 4, nodes_per_game=64, max_batch=4)
    e4.net_create(0)
    with pytest.raises(tb.TakNativeError):                # the DummyNet has nothing to train
        e4.train_begin(64)
    e4.close()
    e6 = tb.Engine(6, 4, nodes_per_game=64, max_batch=4)
    e6.net_create(6)
    with pytest.raises(tb.TakNativeError):                # no weights loaded yet
        e6.train_begin(64)
    e6.net_load_weights(W.random_weights(6, seed=1))
    with pytest.raises(tb.TakNativeError):                # train_chunk before train_begin
        e6.train_chunk(*make_chunk(8, 1))
    e6.train_begin(16)
    with pytest.raises(tb.TakNativeError) as ex:          # chunk larger than reserved
        e6.train_chunk(*make_chunk(17, 1))
    assert ex.value.code == -34
    lp, lz = e6.train_chunk(*make_chunk(1, 2))            # a single position is a valid chunk
    assert np.isfinite([lp, lz]).all()
    e6.train_end()
    with pytest.raises(tb.TakNativeError):
        e6.train_step()
    e6.close()
