"""Shared helpers for the parity tests."""
import numpy as np

import oracle


def splitmix(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def random_positions(n, count, seed=1, half_komi=4, min_ply=2, max_ply=120):
    """Oracle games cut at pseudo-random plies of uniform random playouts (ongoing positions only)."""
    out = []
    gid = 0
    while len(out) < count:
        g = oracle.Game(n, half_komi)
        target = min_ply + splitmix(seed * 7919 + gid) % (max_ply - min_ply)
        ply = 0
        while g.result() == 0 and ply < target:
            moves = g.possible_moves()
            g.play(moves[splitmix(seed * 104729 + gid * 1000003 + ply) % len(moves)])
            ply += 1
        if g.result() == 0:
            out.append(g)
        gid += 1
    return out


def to_tb_state(state):
    """oracle.TakState -> tak_b200.TakState (same POD layout)."""
    import tak_b200 as tb
    return tb.TakState.from_buffer_copy(bytes(state))
