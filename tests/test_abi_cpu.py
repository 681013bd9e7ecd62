"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/taknative.h declares, the
cold host helpers agree with the oracle, and there is no CPU fallback (compute entry points fail loudly)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle
import tak_b200 as tb
from tak_b200 import _lib, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "taknative.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int32_t|const char\*)\s+(\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 40
    lib = tb.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in taknative.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes binding"
    assert sorted(_lib.SYMBOLS) == names
    out = subprocess.check_output(["nm", "-D", "--defined-only", tb.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_struct_layouts_match_header():
    assert C.sizeof(tb.TakState) == 16 + 64 + 64 + 512 + 512 == C.sizeof(oracle.TakState)
    assert C.sizeof(_lib.EngineConfig) == 32
    assert C.sizeof(_lib.SelfplayConfig) == 64
    assert C.sizeof(_lib.SelfplayStats) == 96
    assert C.sizeof(tb.ReplayRecord) == 16 + C.sizeof(tb.TakState) + 512 * 2 + 512 * 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tb.TakNativeError) as ei:
        tb.Engine(5, 4)
    assert ei.value.code == -33 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tak_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import oracle|from oracle)", text, flags=re.M), f
                assert not re.search(r"#include\s+[<\"][^>\"]*oracle", text), f
                assert "liboracle" not in text and "orc_" not in text, f


def test_ptn_round_trip_and_move_index_vs_oracle():
    for n in (3, 4, 5, 6, 7, 8):
        assert tb.policy_size(n) == oracle.policy_size(n)
        assert tb.input_channels(n) == oracle.input_channels(n) == W.input_channels(n)
        g = oracle.Game(n, 0)
        seen = 0
        ply = 0
        while g.result() == 0 and ply < 80:
            moves = g.possible_moves()
            for mv in moves:
                text = oracle.format_move(mv, n)
                assert tb.parse_move(text, n) == mv and tb.format_move(mv, n) == text
                assert tb.move_index(mv, n) == oracle.move_index(mv, n)
                seen += 1
            g.play(moves[(ply * 7919 + n) % len(moves)])
            ply += 1
        assert seen > (150 if n == 3 else 500)


def test_move_index_5_golden(golden_moves_5):
    # reference: alpha-tak/src/search/move_map.rs:51-201
    for i, text in enumerate(golden_moves_5):
        assert tb.move_index(tb.parse_move(text, 5), 5) == i
    with pytest.raises(tb.TakNativeError):
        tb.move_index(tb.parse_move("5a1>11111", 6) | 0, 5)   # not in the 5x5 list


def test_tps_round_trip_vs_oracle(golden):
    t = golden["tps"]
    g = oracle.Game.from_ptn_moves(t["n"], t["moves"])
    st = tb.TakState.from_buffer_copy(bytes(g.state()))
    assert tb.tps_format(st) == t["tps"]            # tak/tests/tps.rs:5-24
    back = tb.tps_parse(t["n"], t["tps"])
    assert back.key() == bytes(oracle.Game.from_tps(t["n"], t["tps"]).state())
    with pytest.raises(tb.TakNativeError):
        tb.tps_parse(6, "x6/x6/x6 1 1")
    with pytest.raises(tb.TakNativeError):
        tb.parse_move("j9", 6)


def test_state_init():
    for n, (s, c) in {3: (10, 0), 4: (15, 0), 5: (21, 1), 6: (30, 1), 7: (40, 2), 8: (50, 2)}.items():
        st = tb.state_init(n, 4)
        assert (st.white_stones, st.white_caps, st.black_stones, st.black_caps) == (s, c, s, c)
        assert st.key() == bytes(oracle.Game(n, 4).state())


def test_weight_blob_layout():
    assert W.blob_size(6) == 5_139_708 and W.blob_size(5) == 7_497_896
    blob = W.random_weights(6, seed=0)
    parts = W.split(blob, 6)
    assert parts["initial_conv.weight"].shape == (128, 92, 3, 3)
    assert parts["policy_conv.weight"].shape == (251, 128, 3, 3)
    assert parts["value_fc.weight"].shape == (1, 4608)
    assert np.array_equal(W.random_weights(6, seed=0), blob) and not np.array_equal(W.random_weights(6, seed=1), blob)


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """The boundary is a C ABI: include/taknative.h compiles as C99 (-pedantic) and a C program links against the
    library and calls its host-side entry points (what a cgo / Rust `extern "C"` consumer relies on)."""
    src = tmp_path / "consumer.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "taknative.h"
int main(void) {
    tak_state_t s;
    uint16_t mv = 0;
    int32_t idx = -1, size = 0;
    char buf[256];
    if (tak_state_init(6, 4, &s) != TAK_OK || s.n != 6 || s.half_komi != 4) return 1;
    if (tak_ptn_parse(6, "3c3>12", &mv) != TAK_OK) return 2;
    if (tak_ptn_format(6, mv, buf, (int32_t)sizeof buf) < 0 || strcmp(buf, "3c3>12") != 0) return 3;
    if (tak_move_index(6, mv, &idx) != TAK_OK || tak_policy_size(6, &size) != TAK_OK || idx < 0 || idx >= size) return 4;
    if (tak_tps_format(&s, buf, (int32_t)sizeof buf) < 0) return 5;
    if (tak_ptn_parse(6, "not a move", &mv) == TAK_OK) return 6;      /* errors are codes, never aborts */
    printf("%d %d %s %d\n", (int)idx, (int)size, buf, (int)sizeof(tak_move_info_t));
    return 0;
}
''')
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(tb.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-ltaknative", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) == tb.move_index(tb.parse_move("3c3>12", 6), 6) and int(out[1]) == 9036
    assert out[2].startswith("x6/x6/x6/x6/x6/x6") and int(out[-1]) == C.sizeof(_lib.MoveInfoRecord)
