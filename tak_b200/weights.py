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


# ---- tch-rs `.model` files: Network::save / Network::load (net6.rs:87-96, net5.rs:95-104) -----------
# `VarStore::save` hands every variable to libtorch's `torch::serialize::OutputArchive` (tch 0.7.2 torch_api.cpp
# `at_save_multi`), i.e. the file is a TorchScript module archive whose PARAMETERS are the variables.  All layers of Net5 /
# Net6 are built on `vs.root()`, so the names are just `weight`, `bias`, `running_mean`, `running_var`; a name that is
# already taken gets the suffix `__<number of variables created so far>` (tch `Path::add`) -- which is what lets the
# creation order be recovered although the file itself is written in HashMap order.  The loader only relies on: that
# suffix rule, the layer creation order of net6.rs:39-57 / net5.rs:39-62, and base names within a layer (it does NOT
# assume whether a conv creates its bias before its weight, or a BatchNorm its running statistics before gamma / beta).
# No `.model` file ships with the reference, so this is parity-unpinned against a real tch file;
# tests/test_weights_cpu.py round-trips through `save_tch_model` under every ordering convention.
_BASES = ("weight", "bias", "running_mean", "running_var")


def _creation_order(names):
    """{name: creation index} from tch's de-duplication suffixes."""
    idx, plain = {}, []
    for nm in names:
        base, sep, suffix = nm.rpartition("__")
        if sep and base in _BASES and suffix.isdigit():
            idx[nm] = int(suffix)
        elif nm in _BASES:
            plain.append(nm)
        else:
            raise ValueError(f"unexpected variable name {nm!r} in a Net5/Net6 VarStore")
    free = sorted(set(range(len(names))) - set(idx.values()))
    if len(free) != len(plain) or len(set(idx.values())) != len(idx):
        raise ValueError("variable names do not follow tch's `name__<count>` de-duplication rule")
    # first use of each name: the conv's two variables come before the BatchNorm's running statistics
    first_use = [n for n in ("weight", "bias") if n in plain] + [n for n in ("running_mean", "running_var") if n in plain]
    if len(first_use) != len(plain):
        raise ValueError("duplicate unsuffixed variable names")
    wb = sorted(n for n in first_use if n in ("weight", "bias"))
    # `weight` / `bias` share the first conv (either order), the running statistics follow in mean, var order
    for nm, i in zip(wb + [n for n in first_use if n.startswith("running")], free):
        idx[nm] = i
    return idx


def blob_from_named_tensors(named: dict, arch: int) -> np.ndarray:
    """The weight blob (`spec(arch)` order) from {tch variable name: array}."""
    order = _creation_order(list(named))
    seq = sorted(named, key=lambda nm: order[nm])
    base_of = lambda nm: nm.rpartition("__")[0] if "__" in nm else nm
    pos, parts = 0, {}

    def take(k, want):
        nonlocal pos
        grp = {base_of(nm): np.asarray(named[nm], dtype=np.float32) for nm in seq[pos:pos + k]}
        if sorted(grp) != sorted(want):
            raise ValueError(f"layer {pos}: expected variables {want}, file has {sorted(grp)}")
        pos += k
        return grp

    sp = spec(arch)
    i = 0
    while i < len(sp):
        name, shape = sp[i]
        layer = name.rsplit(".", 1)[0]
        if ".bn" in name or name.startswith("initial_bn"):
            grp = take(4, _BASES)
            for k in _BASES:
                parts[f"{layer}.{k}"] = grp[k]
            i += 4
        else:
            grp = take(2, ("weight", "bias"))
            parts[f"{layer}.weight"], parts[f"{layer}.bias"] = grp["weight"], grp["bias"]
            i += 2
    if pos != len(seq):
        raise ValueError(f"{len(seq) - pos} variables left over: not a Net{arch} VarStore")
    out = []
    for name, shape in sp:
        a = parts[name]
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: shape {tuple(a.shape)} in the file, Net{arch} needs {tuple(shape)}")
        out.append(a.reshape(-1))
    return np.concatenate(out).astype(np.float32)


def load_tch_model(path: str, arch: int) -> np.ndarray:
    """`Network::load(path)` -> the fp32 weight blob `net_load_weights` takes."""
    import torch
    mod = torch.jit.load(path, map_location="cpu")
    named = {n: p.detach().float().numpy() for n, p in mod.named_parameters()}
    named.update({n: b.detach().float().numpy() for n, b in mod.named_buffers()})
    return blob_from_named_tensors(named, arch)


def tch_variable_names(arch: int, conv_bias_first: bool = True, bn_stats_first: bool = True):
    """[(tch variable name, blob tensor name)] in creation order, as `Net5/Net6::default()` would register them."""
    out, used = [], set()

    def add(base, blob_name):
        nm = base if base not in used else f"{base}__{len(out)}"
        used.add(base)
        out.append((nm, blob_name))

    sp = [name for name, _ in spec(arch)]
    i = 0
    while i < len(sp):
        layer = sp[i].rsplit(".", 1)[0]
        if ".bn" in sp[i] or sp[i].startswith("initial_bn"):
            seq = ("running_mean", "running_var", "weight", "bias") if bn_stats_first else _BASES
            for b in seq:
                add(b, f"{layer}.{b}")
            i += 4
        else:
            for b in (("bias", "weight") if conv_bias_first else ("weight", "bias")):
                add(b, f"{layer}.{b}")
            i += 2
    return out


def save_tch_model(blob: np.ndarray, arch: int, path: str, conv_bias_first: bool = True, bn_stats_first: bool = True,
                   shuffle_seed: int = 0) -> None:
    """`Network::save(path)`: the blob as a TorchScript parameter archive with tch's variable names, written in an
    arbitrary (seeded) order like the HashMap iteration of `VarStore::save`."""
    import torch
    tensors = split(np.asarray(blob, dtype=np.float32), arch)
    names = tch_variable_names(arch, conv_bias_first, bn_stats_first)
    order = np.random.default_rng(shuffle_seed).permutation(len(names))

    class VarStore(torch.nn.Module):
        pass

    m = VarStore()
    for k in order:
        nm, blob_name = names[k]
        trainable = "running_" not in nm
        m.register_parameter(nm, torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(tensors[blob_name])).clone(),
                                                    requires_grad=trainable))
    torch.jit.script(m).save(path)
