"""Weight blobs for Net5 / Net6 in the order `net_load_weights` expects ("weight blob").

The order follows the reference's VarStore creation order (alpha-tak/src/model/net6.rs:39-57, net5.rs:39-62):
initial conv (weight, bias), initial BN (gamma, beta, running_mean, running_var), then per residual block
conv1 (w, b), conv2 (w, b), bn1, bn2, then the policy head (w, b) and the value head (w, b).
"""
from __future__ import annotations

import numpy as np

FILTERS = 128


def stones_caps(n: int):
    return {3: (10, 0), 4: (15, 0), 5: (21, 1), 6: (30, 1), 7: (40, 2), 8: (50, 2)}[n]


def input_channels(n: int) -> int:
    s, c = stones_caps(n)
    return (n + 8) * 2 + 2 + 2 * s + 2 * c


def spec(arch: int):
    """[(name, shape)] of every tensor of the blob, in order."""
    n = arch
    blocks = 16 if arch == 6 else 8
    cin = input_channels(n)
    out = [("initial_conv.weight", (FILTERS, cin, 3, 3)), ("initial_conv.bias", (FILTERS,))]
    out += [(f"initial_bn.{k}", (FILTERS,)) for k in ("weight", "bias", "running_mean", "running_var")]
    for b in range(blocks):
        out += [(f"block{b}.conv1.weight", (FILTERS, FILTERS, 3, 3)), (f"block{b}.conv1.bias", (FILTERS,)),
                (f"block{b}.conv2.weight", (FILTERS, FILTERS, 3, 3)), (f"block{b}.conv2.bias", (FILTERS,))]
        for bn in ("bn1", "bn2"):
            out += [(f"block{b}.{bn}.{k}", (FILTERS,)) for k in ("weight", "bias", "running_mean", "running_var")]
    if arch == 6:
        ch = 3 + 4 * (2 ** 6 - 2)
        out += [("policy_conv.weight", (ch, FILTERS, 3, 3)), ("policy_conv.bias", (ch,))]
    else:
        out += [("policy_fc.weight", (1575, FILTERS * n * n)), ("policy_fc.bias", (1575,))]
    out += [("value_fc.weight", (1, FILTERS * n * n)), ("value_fc.bias", (1,))]
    return out


def blob_size(arch: int) -> int:
    return int(sum(int(np.prod(s)) for _, s in spec(arch)))


def split(blob: np.ndarray, arch: int):
    out, off = {}, 0
    for name, shape in spec(arch):
        k = int(np.prod(shape))
        out[name] = blob[off:off + k].reshape(shape)
        off += k
    assert off == blob.size
    return out


def random_weights(arch: int, seed: int = 0, trained_like: bool = True) -> np.ndarray:
    """Random-init blob with tch-rs' default initialisers: conv / linear weights
    Kaiming-uniform U(+-1/sqrt(fan_in)), conv bias 0, linear bias U(+-1/sqrt(fan_in)), BN gamma U(0,1), beta 0,
    running_mean 0, running_var 1.  With `trained_like` the BN running statistics and biases are perturbed so
    that BN folding is actually exercised (a fresh net has mean 0 / var 1)."""
    rng = np.random.default_rng(seed)
    parts = []
    for name, shape in spec(arch):
        k = int(np.prod(shape))
        if name.endswith("conv.weight") or name.endswith("conv1.weight") or name.endswith("conv2.weight") \
                or name.endswith("fc.weight"):
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            v = rng.uniform(-bound, bound, k)
        elif name.endswith("fc.bias"):
            fan_in = FILTERS * arch * arch
            v = rng.uniform(-1 / np.sqrt(fan_in), 1 / np.sqrt(fan_in), k)
        elif name.endswith(".bias") and ("bn" in name):
            v = rng.uniform(-0.1, 0.1, k) if trained_like else np.zeros(k)
        elif name.endswith(".bias"):
            v = rng.uniform(-0.05, 0.05, k) if trained_like else np.zeros(k)
        elif name.endswith("running_mean"):
            v = rng.uniform(-0.2, 0.2, k) if trained_like else np.zeros(k)
        elif name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, k) if trained_like else np.ones(k)
        elif name.endswith(".weight"):  # BN gamma
            v = rng.uniform(0.0, 1.0, k)
        else:
            raise AssertionError(name)
        parts.append(v.astype(np.float32))
    return np.concatenate(parts)
