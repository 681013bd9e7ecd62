"""The two stand-alone kernel self-tests (tests/cuda/*.cu, built by __graft_entry__.build() into build/) run from pytest,
so that the GPU test run exercises them too:

* conv_selftest: the tcgen05 conv tower against a naive fp32 reference on random strip planes, one layer per launch
  against the fused multi-layer launch (bit-identical across groupings), on the padded strip, the pad-free strip (6x6
  and 5x5) and with a 6-slab first layer;
* wgrad_selftest: the tiled-TMA weight-gradient GEMM (out-of-tile slot groups zero-filled by the TMA unit, split-K over
  tiles) against a one-thread-per-weight reference, for 128 and 96 input channels, both strip pitches, a tile count below
  and above the number of CTAs, and its accumulate mode.
"""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(binary, *args):
    path = os.path.join(ROOT, "build", binary)
    if not os.path.exists(path):   # normally built by __graft_entry__.build(); same command here
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cc = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-I",
                             os.path.join(ROOT, "tak_b200", "csrc"), os.path.join(ROOT, "tests", "cuda", binary + ".cu"),
                             "-o", path], capture_output=True, text=True, timeout=900)
        if cc.returncode != 0:
            pytest.skip(f"{binary} is not built and nvcc could not build it here: {cc.stderr[-300:]}")
    out = subprocess.run([path, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SELFTEST PASSED" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.parametrize("args", [(300, 6, 8, 0), (301, 6, 8, 1), (333, 5, 8, 1), (128, 6, 6, 1)],
                         ids=["padded-6x6", "padfree-6x6", "padfree-5x5", "padfree-6-slabs"])
def test_conv_tower_selftest(args):
    out = run("conv_selftest", *args)
    if args[2] == 8:                       # the fused-vs-separate comparison runs on full 128-channel layers
        assert "identical" in out


@pytest.mark.parametrize("args", [(37, 128, 42), (19, 96, 48), (667, 128, 42), (200, 96, 42)],
                         ids=["few-tiles", "net5-first-layer", "chunk-of-4000", "net6-first-layer"])
def test_wgrad_selftest(args):
    run("wgrad_selftest", *args)
