#!/usr/bin/env python3
"""bench.py -- 6x6 batched self-play (800 rollouts/move, random-init Net6) on N B200s: moves/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--games G] [--rollouts R] [--impl reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W      (N > 1)

A "step" is one searched ply of every concurrent game: forced opening / instant-win scan / Dirichlet noise /
R x (virtual rollout of all G games -> one batched Net6 evaluation -> devirtualise) / pick / replay record /
re-root + play (train/src/self_play.rs:96-262).  `value` = searched plies per second over all ranks with everything
resident in HBM; `e2e` = the same 800-rollout searches driven through the host-buffer C ABI (host game states in,
host visit counts + picked moves out, copies inside the timed region).
`--impl reference` times the CPU restatement of the reference's loop (oracle/ + fp32 PyTorch Net6, all host threads)
on a bounded sample of the same workload; it is the only place bench.py executes oracle/ besides `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "self-play MCTS moves/sec (6x6, 800 rollouts)"
UNIT = "moves/s"
FLOP_PER_EVAL_NET6 = 368_197_632


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"], "hbm": p["hbm_gbs"],
                "source": "MEASURED_PEAKS.json (measured)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
               "samples": len(sm)}
        if pw:
            out["power_w"] = float(np.median(pw))
            try:
                lim = subprocess.run(["nvidia-smi", "--query-gpu=power.limit", "--format=csv,noheader,nounits", "-i",
                                      str(self.device)], capture_output=True, text=True, timeout=10).stdout.strip()
                out["power_limit_w"] = float(lim)
            except Exception:
                pass
        return out


# ---------------------------------------------------------------------------------------------------------------
# CPU restatement of the reference loop (oracle + fp32 torch Net6): cpu_baseline and --impl reference
# ---------------------------------------------------------------------------------------------------------------
def cpu_selfplay_sample(rollouts_sample: int, full_rollouts: int, workers: int = 32, seed: int = 0):
    """32 lock-step games (WORKERS, self_play.rs:94), one leaf per tree per step, one batched policy_eval per step.
    Runs `rollouts_sample` of the `full_rollouts` rollouts of one ply and scales."""
    import torch

    import oracle
    from oracle.net_ref import RefNet
    from tak_b200 import weights as W

    oracle.use_native_build()     # -O3 -march=native on this box's CPU, as .cargo/config.toml:2 builds the reference
    n = 6
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    net = RefNet(6, W.random_weights(6, seed=seed), device="cpu")
    games, searches = [], []
    for i in range(workers):
        g = oracle.Game.with_komi(n, 2)
        g.play("a1")
        g.play("a6" if i % 2 else "f6")
        games.append(g)
        searches.append(oracle.Search(n))
    cores = torch.get_num_threads()
    t0 = time.perf_counter()
    evals = 0
    for _ in range(rollouts_sample):
        pend = [i for i in range(workers) if searches[i].virtual_rollout(games[i]) == 0]
        if not pend:
            continue
        x = np.stack([oracle.Game.from_state(searches[i].pending_state(0)).repr() for i in pend])
        pol, val, _ = net.forward_mcts(torch.from_numpy(x))
        pol, val = pol.numpy(), val.numpy()
        for j, i in enumerate(pend):
            searches[i].devirtualize(pol[j], float(val[j]))
        evals += len(pend)
    dt = time.perf_counter() - t0
    moves_per_s = workers / (dt * full_rollouts / rollouts_sample)
    return {"value": moves_per_s, "seconds": dt, "cores": cores, "evals": evals,
            "sample": f"{workers} lock-step games x {rollouts_sample} of {full_rollouts} rollouts of one ply "
                      f"(oracle MCTS + fp32 PyTorch Net6, batch {workers}), scaled by {full_rollouts}/{rollouts_sample}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = max(8, min(args.rollouts, 40))
    cpu_selfplay_sample(4, args.rollouts)  # warm-up (thread pools, oneDNN primitives)
    for _ in range(max(0, args.warmup - 1)):
        cpu_selfplay_sample(4, args.rollouts)
    t0 = time.perf_counter()
    res = [cpu_selfplay_sample(sample, args.rollouts) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    value = float(np.mean([r["value"] for r in res]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "6x6 batched self-play, 800 rollouts/move, random-init Net6 (CPU restatement of the "
                               "reference loop: the Rust reference cannot be built here)",
                   "games": 32, "rollouts": args.rollouts},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res[0]["cores"], "kind": "port",
                         "sample": res[0]["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def in_threads(fns):
    """Run the callables concurrently (ctypes releases the GIL inside the C ABI) and return their results in order."""
    out = [None] * len(fns)
    err = []

    def wrap(i, f):
        try:
            out[i] = f()
        except BaseException as ex:  # noqa: BLE001 - re-raised below
            err.append(ex)

    ts = [threading.Thread(target=wrap, args=(i, f)) for i, f in enumerate(fns)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if err:
        raise err[0]
    return out


PERFT6_D5 = 1_253_506_520  # tak/tests/perft.rs:98 (the value the reference keeps commented out)
PERFT5_D4 = 2_999_784      # tak/tests/perft.rs:64
STATE_BYTES = {3: 160, 4: 192, 5: 288, 6: 384, 7: 896, 8: 1152}


def bytes_per_node(n, branching):
    """S (child state written) + S/b (parent read, amortised over its b children) + 2 B (move)."""
    S = STATE_BYTES[n]
    return S + S / max(branching, 1.0) + 2


def movegen_mnodes(eng, world, rank, dev, pk, red):
    """The metric's second half: movegen + play + result throughput as perft(5) of the 6x6 opening position through the
    C ABI (tak_perft), breadth-first on the device.  Two DIFFERENT rates come out of one run and are reported apart:
      * counted: perft's node count / time -- 99 % of the 1.25e9 nodes are depth-5 leaves that perf_count only COUNTS
        (perft.rs:6-7: `possible_moves().len()`), in closed form from the tail of their depth-4 parent;
      * materialised: positions actually generated, applied, classified and written to HBM (depth <= 4: 13.7 M states).
        The roofline is stated on the largest expansion (depth 3 -> 4, 13.59 M children) with SURVEY 8(d)'s bytes.
    With N ranks the 1 260 positions two plies below the root are dealt round-robin (tak_perft_multi expands a rank's share
    as one frontier) and the counts are summed with one all-reduce."""
    depth = 5
    eng.reset(0, 1, 0)                       # slot 0 of this rank's engine (self-play is over) holds the opening
    root = eng.download([0])[0]
    eng.perft(root, depth)                   # warm-up at full depth: the frontier arenas are allocated here
    best = None
    for _ in range(3):
        if world == 1:
            nodes = eng.perft(root, depth)
        else:
            # every rank builds the depth-2 frontier (1 260 positions, host-driven, untimed), takes every world-th
            # position and expands its share in ONE breadth-first perft of the remaining depth
            if best is None:
                front, ended = eng.frontier(root, 2)
                mine = front[rank::world]
                eng.perft_multi(mine, depth - 2)     # warm-up: arenas sized for this share
            nodes = eng.perft_multi(mine, depth - 2) + (ended if rank == 0 else 0)
        prof = eng.perft_profile()
        if best is None or prof["ms"] < best[1]["ms"]:
            best = (nodes, prof)
    nodes, prof = best
    total = int(red.sum(nodes))
    t_max = red.max(prof["ms"])
    mat_total = red.sum(prof["materialised"])
    # the largest expansion of this rank: children / parents = branching; bytes by the 8(d) formula
    parents = max(1, prof["materialised"] - prof["top_children"]) if world == 1 else None
    b = prof["top_children"] / parents if parents else 100.0
    bpn = bytes_per_node(6, b)
    gbs = prof["top_children"] * bpn / (prof["top_ms"] * 1e-3) / 1e9 if prof["top_ms"] else 0.0
    return {"value": total / (t_max * 1e-3) / 1e6 if t_max else None, "unit": "Mnodes/s",
            "what": "COUNTED rate: perft node count / device time of the whole tak_perft call (best of 3)",
            "workload": "6x6 perft depth 5 from the opening via tak_perft (movegen + play + result, bit-exact count)"
                        + ("" if world == 1 else f"; the 1260 positions two plies down are dealt round-robin over {world} ranks "
                           "(that shallow frontier is built on the host, untimed), each rank's share is one tak_perft_multi"),
            "nodes": total, "exact": total == PERFT6_D5, "ms": t_max,
            "materialised": {"states": int(mat_total), "gstates_per_s": mat_total / (t_max * 1e-3) / 1e9 if t_max else None,
                             "what": "positions generated + applied + classified + written to HBM (depth <= 4), over the "
                                     "time of the WHOLE call incl. scans, move lists and host round trips for the "
                                     "frontier sizes; the depth-5 leaves are only counted"},
            "roofline": {"bound": "hbm", "kernel": "k_perft_moves + k_perft_apply, depth 3 -> 4 expansion (rank 0)",
                         "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"] if pk["hbm"] else None,
                         "children": prof["top_children"], "ms": prof["top_ms"],
                         "gstates_per_s": prof["top_children"] / (prof["top_ms"] * 1e-3) / 1e9 if prof["top_ms"] else None,
                         "bytes_per_node": bpn, "branching": b,
                         "formula": "S + S/b + 2 per materialised child (SURVEY.md 8d), S = 384",
                         # dram__bytes_read.sum + dram__bytes_write.sum of k_perft_apply<6,true> for the same expansion
                         # (ncu --set full, profiles/r02_perft_apply_ncu.txt): 5.16 GB written + 0.08 GB read
                         "traffic": 5_240_076_000,
                         "note": "ceiling by the same formula: peak / bytes_per_node = "
                                 f"{pk['hbm'] / bpn:.1f} G states/s"}}


def perft5_rate(local):
    """configs[0] on the device: 5x5 perft(4) from the opening (tak/tests/perft.rs:57-64), the reference's CPU-runnable
    case.  43 945 states are materialised, so the call is dominated by launch latency and the host reading back the
    frontier sizes, not by bandwidth."""
    import tak_b200 as tb
    eng = tb.Engine(5, 4, device=local, nodes_per_game=64)
    eng.reset(0, 1, 0)
    root = eng.download([0])[0]
    eng.perft(root, 4)
    best = None
    for _ in range(5):
        nodes = eng.perft(root, 4)
        prof = eng.perft_profile()
        if best is None or prof["ms"] < best[1]["ms"]:
            best = (nodes, prof)
    eng.close()
    nodes, prof = best
    return {"value": nodes / (prof["ms"] * 1e-3) / 1e6, "unit": "Mnodes/s", "nodes": nodes, "exact": nodes == PERFT5_D4,
            "ms": prof["ms"], "materialised_states": prof["materialised"], "launches": prof["launches"],
            "workload": "5x5 perft depth 4 from the opening via tak_perft (configs[0]); launch/latency-bound at this size"}


def stress8_rates(local, world, rank, red, pk):
    """configs[4], 8x8 stress (deep stacks, long spreads, u128 columns), every rank on its own games (weak scaling):
      A. tak_playouts: uniform-random playouts to termination from positions cut at ply 60..200 of other random playouts;
         nodes = plies played (one generate + select + apply + result each); the game state stays in registers for the
         whole playout, so this is issue/latency-bound, not HBM-bound (HBM sees one load and one store per game).
      B. MCTS with the DummyNet prior (alpha-tak/src/search/tests.rs:29-34), 800 rollouts per position, one launch for
         the whole search (nothing to evaluate, so nothing forces a kernel boundary)."""
    import tak_b200 as tb
    G = 148 * 32                          # 800 rollouts x ~200 children per expansion: 262 144 nodes per game and half
    eng = tb.Engine(8, G, device=local, nodes_per_game=1 << 18, max_batch=64)
    base = rank * G
    eng.reset(0, G, 4)
    eng.playouts(0, G, 0x8A8, 60, 141, game_id_base=base)        # cut positions, ply 60..200 (untimed)
    ids = np.arange(G, dtype=np.int32)
    start = eng.download(ids[:64])
    deep = max(max(s.height) for s in start)
    # B first (it leaves the positions unchanged)
    eng.net_create(0)
    eng.tree_reset(ids)
    eng.rollouts(ids, 8)                                         # warm-up
    eng.tree_reset(ids)
    eng.sync()
    t0 = time.perf_counter()
    eng.rollouts(ids, 800)
    dt_b = red.max(time.perf_counter() - t0)
    root_visits = eng.root(0)[0]
    eng.close()
    # A: its own engine -- no search trees, so tens of thousands of games fit and fill the SMs
    GA = 148 * 512
    eng = tb.Engine(8, GA, device=local, nodes_per_game=64, max_batch=64)
    base = rank * GA
    eng.reset(0, GA, 4)
    eng.playouts(0, GA, 0x8A8, 60, 141, game_id_base=base)       # cut positions, ply 60..200 (untimed)
    plies, res, tot = eng.playouts(0, GA, 0x8A9, 100_000, 0, game_id_base=base)
    ms_a = red.max(tot["ms"])
    plies_all = red.sum(tot["plies"])
    gen_all = red.sum(tot["generated"])
    eng.close()
    return {
        "playouts": {"value": plies_all / (ms_a * 1e-3) / 1e6, "unit": "Mnodes/s", "nodes": int(plies_all),
                     "moves_generated_per_s": gen_all / (ms_a * 1e-3), "ms": ms_a, "games": GA * world,
                     "finished": int((res != 0).sum()), "deepest_start_stack": int(deep),
                     "workload": "8x8 random playouts to termination from ply-60..200 positions (tak_playouts): nodes = "
                                 "plies = positions generated + applied + classified; state register-resident, "
                                 "issue/latency-bound (no HBM roofline: one 1152-byte load and store per game)"},
        "mcts_dummy": {"value": G * world / dt_b, "unit": "moves/s", "rollouts_per_s": 800.0 * G * world / dt_b,
                       "seconds": dt_b, "games": G * world, "rollouts": 800, "root_visits": int(root_visits),
                       "workload": "8x8 MCTS, DummyNet prior, 800 rollouts per position on ply-60..200 positions "
                                   "(mcts_rollouts: select/expand/backup only, ONE launch per search), host-timed"},
    }


def selfplay5_rate(local, world, rank, red, steps, warmup, blob5):
    """configs[1]: 5x5 self-play, random-init Net5, 800 rollouts/move, Dirichlet noise off, always exploit, 1 B200 (every
    rank its own games when N > 1)."""
    import tak_b200 as tb
    G = 148 * 6 * tb.boards_per_tile(5)   # 6 conv tiles per SM (10 boards per tile on the pad-free 5x5 strip)
    eng = tb.Engine(5, G, device=local, nodes_per_game=1 << 17, max_batch=G)
    eng.net_create(5)
    eng.net_load_weights(blob5)
    eng.selfplay_begin(rollouts=800, half_komi=4, instant_win=1, exploit_ply=0, noise_ply=0, seed=0x55,
                       game_id_base=rank * G)
    for _ in range(warmup):
        eng.selfplay_step(1)
        eng.selfplay_drain()
    ms, plies, evals, launches = 0.0, 0, 0, 0
    for _ in range(steps):
        st = eng.selfplay_step(1)
        ms += st.device_ms
        plies += st.plies_played
        evals += st.evals
        launches += st.kernel_launches
        eng.selfplay_drain()
    eng.close()
    t = red.max(ms)
    return {"value": red.sum(plies) / (t * 1e-3), "unit": "moves/s", "ms_per_step": t / steps, "steps": steps,
            "games": G * world, "evals_per_s": red.sum(evals) / (t * 1e-3), "launches_per_ply": launches / steps,
            "frac_of_tensor_peak": red.sum(evals) / world / (t * 1e-3) * 132_198_400 / (pk_sustained() * 1e12),
            "workload": "5x5 batched self-play, Net5 (random init), 800 rollouts/move, noise off, always exploit, "
                        "instant-win on (configs[1])"}


def late_game_rate(local, blob, steps):
    """The headline steps start from the forced opening, so they cover plies 2..30.  This leg runs the same self-play
    configuration from MID-GAME positions (every game advanced by 10..70 uniform-random plies with tak_playouts first):
    more pieces on the board, taller stacks, wider move lists, terminal rollouts that need no evaluation."""
    import tak_b200 as tb
    bpt = tb.boards_per_tile(6)
    G = 148 * (4 if bpt == 7 else 6) * bpt
    eng = tb.Engine(6, G, device=local, nodes_per_game=1 << 18, max_batch=G)
    eng.net_create(6)
    eng.net_load_weights(blob)
    eng.reset(0, G, 4)
    plies, res, _ = eng.playouts(0, G, 0x1A7E, 10, 61)
    for i in np.nonzero(res)[0]:
        eng.reset(int(i), 1, 4)                                            # a finished playout restarts from the opening
    start = eng.download(np.arange(256, dtype=np.int32))
    eng.selfplay_begin(rollouts=800, half_komi=4, instant_win=1, exploit_ply=40, noise_ply=80, noise_alpha=0.2,
                       noise_ratio=0.3, seed=0x7A4C, keep_positions=True)
    eng.selfplay_step(1)
    eng.selfplay_drain()
    ms, plies_played, evals, rollouts, done_games = 0.0, 0, 0, 0, 0
    for _ in range(steps):
        st = eng.selfplay_step(1)
        ms += st.device_ms
        plies_played += st.plies_played
        evals += st.evals
        rollouts += st.rollouts
        done_games += st.games_completed
        eng.selfplay_drain()
    eng.close()
    return {"value": plies_played / (ms * 1e-3), "unit": "moves/s", "games": G, "replicas": 1, "steps": steps,
            "mean_start_ply": float(np.mean([s.ply for s in start])), "max_start_ply": int(max(s.ply for s in start)),
            "evals_per_rollout": evals / max(1, rollouts), "games_completed": int(done_games),
            "workload": "6x6 self-play as in the headline, but every game starts 10..70 random plies into the game "
                        "(one engine replica; throughput does not depend on the number of replicas, so it "
                        "compares directly with the headline value)"}


def pk_sustained():
    return peaks()["bf16_sustained"]


class Reducer:
    """max / sum over ranks of host values: through the engine's NCCL communicator (C ABI) when there is one."""

    def __init__(self, comm):
        self.comm = comm

    def max(self, x):
        return float(self.comm.max_f64([x])[0]) if self.comm else float(x)

    def sum(self, x):
        if not self.comm:
            return float(x)
        if float(x).is_integer() and x >= 0:          # counts: exact
            return float(self.comm.sum_u64([int(x)])[0])
        return self.comm.sum_u64([int(round(x * 1e6))])[0] / 1e6   # anything else in micro-units


def augment_rate(eng, recs, dev, pk):
    """examples_to_tensors (alpha-tak/src/example.rs:63-78) on replay records of this run, outputs left in HBM."""
    import torch

    import tak_b200 as tb
    from tak_b200._lib import check

    if not recs:
        return None, None
    k = len(recs)
    arr = (tb.ReplayRecord * k)(*recs)
    c, p, n = tb.input_channels(6), tb.policy_size(6), 6
    inputs = torch.empty((8 * k, c, n, n), dtype=torch.float32, device=dev)
    pi = torch.empty((8 * k, p), dtype=torch.float32, device=dev)
    z = torch.empty(8 * k, dtype=torch.float32, device=dev)
    fp = C.POINTER(C.c_float)
    ptr = lambda t: C.cast(t.data_ptr(), fp)
    call = lambda: check(eng.lib.examples_to_tensors(eng._h, arr, k, ptr(inputs), ptr(pi), ptr(z), 1))
    call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    out_bytes = 8 * k * (c * n * n + p + 1) * 4 + 8 * k * p * 4   # tensors written + the zero fill of pi
    return {"value": k / dt, "unit": "examples/s", "examples": k, "rows_out": 8 * k, "ms": 1e3 * dt,
            "what": "host replay records -> H2D -> 8 symmetries x (game_repr, pi, z) in HBM, host-timed incl. the copy",
            "hbm_write_gbs": out_bytes / dt / 1e9, "hbm_peak_gbs": pk["hbm"]}, (inputs, pi, z)


def interactive_rates(local, blob, states):
    """The `analysis` / `playtak` / `pit` / reference-`train` regime (small batches): latency of Network::policy_eval
    through the host-buffer ABI (alpha-tak/src/model/network.rs:34), the rollouts/s of a single `Player` with batch 32
    (analysis/src/main.rs prints the same figure as nps), and self-play at the reference's own size: WORKERS = 32
    lock-step games x 800 rollouts (train/src/self_play.rs:94) -- the regime where launch latency, not throughput, rules."""
    import tak_b200 as tb

    # a search of one game for a second grows a tree of millions of nodes: its own engine with a deep node pool
    eng = tb.Engine(6, 2, device=local, nodes_per_game=1 << 23, max_batch=256)
    eng.net_create(6)
    eng.net_load_weights(blob)
    out = {"policy_eval_ms": {}, "what": "net_policy_eval(host states -> full softmax [b, 9036] + value in pinned host buffers "
                                         "(tak_host_alloc; _pageable: ordinary numpy arrays)), median of 20 calls; Player(batch 32).rollout() on one game for 1 s; "
                                         "selfplay_step on 32 games x 800 rollouts"}
    out["policy_eval_ms_pageable"] = {}
    for b in (1, 32, 256):
        st = (tb.TakState * b)(*states[:b])            # the caller's games as one POD array
        pinned = (eng.pinned_array((b, eng.policy_size)), eng.pinned_array((b,)))   # result buffers from tak_host_alloc
        for key, buf in (("policy_eval_ms", pinned), ("policy_eval_ms_pageable", None)):
            eng.policy_eval(st, out=buf)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter()
                eng.policy_eval(st, out=buf)
                ts.append(time.perf_counter() - t0)
            out[key][str(b)] = 1e3 * float(np.median(ts))
    gid, batch = 0, 32
    eng.reset(gid, 1, 4)
    pl = tb.Player(eng, gid, batch)
    for _ in range(5):
        pl.rollout()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < 1.0 and n * batch < 60_000:
        pl.rollout(8)                       # 8 Player::rollout calls fused into one mcts_player_rollouts
        n += 8
    out["player_nps"] = n * batch / (time.perf_counter() - t0)
    out["player_batch"] = batch
    eng.close()
    # the reference's own self-play size
    eng = tb.Engine(6, 32, device=local, nodes_per_game=1 << 18, max_batch=32)
    eng.net_create(6)
    eng.net_load_weights(blob)
    eng.selfplay_begin(rollouts=800, half_komi=4, instant_win=1, exploit_ply=40, noise_ply=80, noise_alpha=0.2,
                       noise_ratio=0.3, seed=0x7A4B)
    eng.selfplay_step(1)
    t0 = time.perf_counter()
    st = eng.selfplay_step(3)
    dt = time.perf_counter() - t0
    out["moves_per_s_32_games"] = st.plies_played / dt
    out["ms_per_rollout_step_32_games"] = 1e3 * dt / (3 * 801)
    out["launches_per_ply_32_games"] = st.kernel_launches / 3
    eng.close()
    return out


def train_rate(eng, tensors, world, comm, red, pk, host_recs=None):
    """Network::train_inner + Adam (next row N1, network.rs:37-97) on the augmented examples of this run: chunks of
    500 examples x 8 symmetries = 4000 positions (CHUNK_SIZE, network.rs:19), inputs resident in HBM; with N ranks every
    rank trains its own chunks and the fp32 gradient blob is all-reduced over NCCL (net_train_allreduce, on the engine's
    stream) before the Adam step."""
    import torch

    from tak_b200 import weights as W

    inputs, pi, z = tensors
    B = min(4000, inputs.shape[0])
    x, p, zz = inputs[:B].contiguous(), pi[:B].contiguous(), z[:B].contiguous()
    eng.train_begin(B)
    for _ in range(2):
        eng.train_chunk(x, p, zz)
    chunks = 4
    ms, loss = [], None
    for _ in range(chunks):
        loss = eng.train_chunk(x, p, zz)
        ms.append(eng.train_stats()["ms_last_chunk"])
    # end to end through the public API with HOST examples: replay records -> examples_to_tensors (H2D + 8 symmetries
    # on the device) -> train_chunk -> losses back on the host, wall clock
    e2e = None
    if host_recs and len(host_recs) * 8 >= B:
        recs = host_recs[:B // 8]
        eng.train_chunk(*eng.examples_to_tensors(recs, on_device=True))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            eng.train_chunk(*eng.examples_to_tensors(recs, on_device=True))
        torch.cuda.synchronize()
        dt = red.max((time.perf_counter() - t0) / reps)
        e2e = {"value": world * B / dt, "unit": "positions/s", "ms_per_chunk": 1e3 * dt,
               "h2d_bytes_per_step": len(recs) * C.sizeof(type(recs[0])), "d2h_bytes_per_step": 8,
               "what": "host replay records -> examples_to_tensors(on_device) -> net_train_chunk -> (loss_p, loss_z)"}
    t_ar = 0.0
    if comm:                                  # the first call builds NCCL's rings / buffers for this size: untimed
        comm.allreduce_gradients()
        eng.sync()
        t0 = time.perf_counter()
        comm.allreduce_gradients()
        eng.sync()
        t_ar = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.train_step(1e-4, 1e-4)
    eng.sync()
    t_step = time.perf_counter() - t0
    ms_chunk = float(np.mean(ms))
    ms_max = red.max(ms_chunk)
    flop = 3.0 * FLOP_PER_EVAL_NET6 * B          # forward + dgrad + wgrad
    out = {"value": world * B / (ms_max * 1e-3), "unit": "positions/s", "positions_per_chunk": B,
           "ms_per_chunk": ms_max, "allreduce_ms": 1e3 * t_ar, "adam_step_ms": 1e3 * t_step,
           "grad_bytes": int(W.blob_size(6)) * 4, "loss_p": loss[0], "loss_z": loss[1],
           "tflops": flop / (ms_max * 1e-3) / 1e12, "tensor_peak": pk["bf16_sustained"],
           "frac_of_tensor_peak": flop / (ms_max * 1e-3) / 1e12 / pk["bf16_sustained"], "e2e": e2e,
           "what": "train_inner on 4000 augmented positions (forward_training + loss + backward, CUDA events on the engine "
                   "stream), then net_train_allreduce (NCCL on the engine stream) and the Adam step"}
    eng.train_end()
    return out


def run_b200(args):
    import torch

    import tak_b200 as tb
    from tak_b200 import comm as tc
    from tak_b200 import weights as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        # torch.distributed is plumbing only (rendezvous: ships the NCCL unique ids, host barrier); every exchange of
        # data -- weights, replay, gradients, the reductions of the timings -- goes through the engine's own NCCL
        # communicator behind the C ABI (tak_comm_init, net_broadcast_weights, selfplay_gather_replay, ...)
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    dev = torch.device("cuda", local)

    # Games never interact, so a GPU's games are split over `replicas` independent engines (own stream, search trees and
    # network replica) driven by one host thread each.
    E = max(1, args.replicas)
    # games per GPU and replica: a whole number of conv tiles per SM -- 148 SMs x 4 tiles x 7 boards = 4 144 on the pad-free
    # strip (148 x 6 x 6 = 5 328 on the padded one).  Measured on one box, games per GPU -> moves/s: 4 144 -> 3 890,
    # 6 216 -> 3 859, 8 288 -> 3 875-3 882, 10 360 -> 3 821, 12 432 -> 3 824: flat within 2 %, the smaller node pools
    # (104 GB for 8 288 games) cost a little less HBM traffic and power
    bpt = tb.boards_per_tile(6)
    per_replica = 148 * (4 if bpt == 7 else 6) * bpt
    G = args.games if args.games else per_replica * E
    Gr = G // E
    G = Gr * E
    R, n = args.rollouts, 6
    pk = peaks()
    elems = W.blob_size(6)
    engines = []
    for r in range(E):
        eng = tb.Engine(n, Gr, device=local, nodes_per_game=args.nodes_per_game, max_batch=Gr)
        eng.net_create(6)
        engines.append(eng)
    comms = [None] * E
    if world > 1:
        uids = [tc.unique_id() if rank == 0 else None for _ in range(E)]
        dist.broadcast_object_list(uids, src=0)
        comms = [tc.Comm(engines[r], uids[r], rank, world) for r in range(E)]
    red = Reducer(comms[0])

    # weights: rank 0 draws them; NCCL (net_broadcast_weights, on each replica's stream) hands the fp32 blob to every
    # rank, which folds BatchNorm and packs the tensor-core operand images of its own replica
    blob = W.random_weights(6, seed=0) if rank == 0 else None
    comm_stats = {}
    if world > 1:
        for r in range(E):                                  # first use builds NCCL's channels: untimed
            comms[r].broadcast_weights(blob, root=0)
        t0 = time.perf_counter()
        for r in range(E):
            comms[r].broadcast_weights(blob, root=0)
        comm_stats["broadcast_ms"] = 1e3 * red.max(time.perf_counter() - t0) / E
        comm_stats["broadcast_bytes"] = elems * 4
        comm_stats["broadcast_what"] = ("net_broadcast_weights: H2D on the root, ncclBroadcast of the fp32 blob, D2H, "
                                        "BatchNorm folding + operand packing on every rank (per replica)")
    else:
        for eng in engines:
            eng.net_load_weights(blob)
    if blob is None:
        blob = W.random_weights(6, seed=0)                  # same seed: only used by the CPU-side sub-benchmarks below

    def barrier():
        torch.cuda.synchronize()
        for eng in engines:
            eng.sync()
        if dist:
            dist.barrier()

    # the conv tower timed ALONE on a cool GPU (3 launches, CUDA events) -> compared with the burst peak below
    engines[0].reset(0, Gr, 4)
    prof_alone = engines[0].net_forward_profile(0, Gr, 3)

    # ---------------- movegen Mnodes/s: 6x6 perft(5) from the opening, root moves sharded over ranks ----------------
    # (before the self-play steps: afterwards the GPU sits at its power-capped clock for a while)
    movegen = movegen_mnodes(engines[0], world, rank, dev, pk, red)

    # ---------------- device-resident self-play: `value` ----------------
    for r, eng in enumerate(engines):
        eng.selfplay_begin(rollouts=R, half_komi=4, instant_win=1, exploit_ply=40, noise_ply=80, noise_alpha=0.2,
                           noise_ratio=0.3, seed=0x7A4B, game_id_base=(rank * E + r) * Gr)

    def warm(eng):
        for _ in range(args.warmup):
            eng.selfplay_step(1)
            eng.selfplay_drain()

    in_threads([lambda eng=eng: warm(eng) for eng in engines])

    def timed(eng, steps):
        """`steps` searched plies of every game of this replica; device time = CUDA events on the replica's stream."""
        tot = {"ms": 0.0, "launches": 0, "evals": 0, "plies": 0, "done": 0, "recs": []}
        left = steps
        while left > 0:
            k = min(left, 4)                       # the replay ring holds 16 records per game between drains
            st = eng.selfplay_step(k)
            tot["ms"] += st.device_ms
            tot["launches"] += st.kernel_launches
            tot["evals"] += st.evals
            tot["plies"] += st.plies_played
            tot["done"] += st.games_completed
            left -= k
            tot["recs"] += eng.selfplay_drain()
        return tot

    # The trainer publishes a new network while self-play runs (train/src/main.rs:101-105,120): half way through the timed
    # plies every rank takes a weight refresh over NCCL.  Its time is reported (comm.refresh_ms) and is part of the wall
    # clock of the region; `value` counts the searched plies over the device time of the self-play steps themselves.
    half = args.steps // 2 if (world > 1 and args.steps >= 2) else 0
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    res = in_threads([lambda eng=eng: timed(eng, args.steps - half) for eng in engines])
    if half:
        t0 = time.perf_counter()
        for r in range(E):
            comms[r].broadcast_weights(blob if rank == 0 else None, root=0)
        comm_stats["refresh_ms"] = 1e3 * (time.perf_counter() - t0) / E
        res2 = in_threads([lambda eng=eng: timed(eng, half) for eng in engines])
        for a, b2 in zip(res, res2):
            for k in a:
                a[k] += b2[k]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop()
    dev_ms = max(r["ms"] for r in res)                  # replicas run concurrently: the GPU's time is the longest span
    launches = sum(r["launches"] for r in res)
    evals = sum(r["evals"] for r in res)
    plies = sum(r["plies"] for r in res)
    games_done = sum(r["done"] for r in res)
    recs = sum((r["recs"] for r in res), [])
    t_max = red.max(max(dev_ms, 0.0))                   # device time (CUDA events on the engine streams), max over ranks
    total_plies = red.sum(float(plies))
    value = total_plies / (t_max / 1e3)
    total_launches = int(red.sum(float(launches)))
    wall_max = red.max(wall_ms)
    if "refresh_ms" in comm_stats:
        comm_stats["refresh_ms"] = red.max(comm_stats["refresh_ms"])
    replay_bytes = len(recs) * C.sizeof(tb.ReplayRecord)

    # ---------------- end to end through the host-buffer ABI: `e2e` ----------------
    ids = np.arange(Gr, dtype=np.int32)
    host_states = [eng.download(ids) for eng in engines]  # the positions self-play reached, now in host memory
    state_bytes = STATE_BYTES[n]
    h2d = G * state_bytes + G * 4
    stride = 256
    d2h = G * stride * 6 + G * 4 + G * 2
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_run(i, steps):
        eng = engines[i]
        for _ in range(steps):
            eng.upload(ids, host_states[i])        # H2D: packed game states from pinned staging
            eng.tree_reset(ids)
            eng.rollouts(ids, R)                   # select -> encode -> Net6 -> backup, R times
            mv, vis, cnt = eng.children_batch(ids, stride)   # D2H: improved policy (visit counts) of every root
            picks = eng.pick_move(ids)             # D2H: the moves to play
        return mv, vis, cnt

    in_threads([lambda i=i: e2e_run(i, 1) for i in range(E)])
    barrier()
    t0 = time.perf_counter()
    e2e_out = in_threads([lambda i=i: e2e_run(i, e2e_steps) for i in range(E)])
    barrier()
    e2e_t = red.max(time.perf_counter() - t0)
    e2e_value = world * G * e2e_steps / e2e_t

    # ---------------- roofline of the dominant kernel (conv3x3_tc3_kernel = the whole conv tower), measured live -------
    # Timed right after the self-play steps, 10 launches back to back on the hot, power-capped GPU: the conditions of the
    # timed region (per-launch events INSIDE the region would also count the other replica's kernels sharing the SMs), so
    # the denominator is the SUSTAINED peak; the same kernel timed alone before the run goes against the BURST peak.
    prof = engines[0].net_forward_profile(0, Gr, 10)
    value_fc_flop = Gr * (2.0 * 128 * 36)            # all but the value FC runs in the conv kernel
    conv_flop = prof["flop"] - value_fc_flop
    achieved = conv_flop / (prof["ms_conv"] * 1e-3) / 1e12
    alone = (prof_alone["flop"] - value_fc_flop) / (prof_alone["ms_conv"] * 1e-3) / 1e12
    roofline = {
        "bound": "tensor", "kernel": "conv3x3_tc3_kernel", "achieved": achieved, "peak": pk["bf16_sustained"],
        "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
        # dram__bytes_read.sum + dram__bytes_write.sum of one tower launch over 5180 boards (ncu --set full,
        # profiles/r02_conv_tc3_padfree_ncu_full.txt), scaled to this launch's boards: logits + write-backs of activations
        "traffic": 678_200_320 * Gr / 5180,
        "peak_kind": "sustained bf16 (kernel timed in a back-to-back loop under the step's power cap, CUDA events around "
                     "each launch), " + pk["source"],
        "avg_launch_us": 1e3 * prof["ms_conv"] / prof["conv_launches"],
        "algorithmic_flop_per_launch": conv_flop / prof["conv_launches"],
        "boards_per_launch": Gr,
        "alone": {"achieved": alone, "peak": pk["bf16_burst"], "frac": alone / pk["bf16_burst"],
                  "avg_launch_us": 1e3 * prof_alone["ms_conv"] / prof_alone["conv_launches"],
                  "peak_kind": "burst bf16: the same launch timed alone on the cool GPU before the run"},
        "step_frac_sustained": (value / world) * R * FLOP_PER_EVAL_NET6 / (pk["bf16_sustained"] * 1e12),
        "forward_ms": prof["ms_forward"], "conv_share_of_forward": prof["ms_conv"] / prof["ms_forward"],
    }

    # ---------------- replay augmentation (next row N2): Example::to_tensors x 8 symmetries on the device ------------
    mv0, vis0, cnt0 = e2e_out[0]
    aug_recs = []
    for i in range(min(2048, Gr)):             # the e2e searches' (position, improved policy) pairs as Examples
        r = tb.ReplayRecord()
        r.state, r.result, r.n_children = host_states[0][i], 1.0 - 2.0 * (i & 1), int(cnt0[i])
        C.memmove(r.moves, mv0[i].ctypes.data, 2 * r.n_children)
        C.memmove(r.visits, vis0[i].ctypes.data, 4 * r.n_children)
        aug_recs.append(r)
    augment, aug_tensors = augment_rate(engines[0], aug_recs, dev, pk)
    # replay gather over NCCL (selfplay_gather_replay): every rank contributes the records above (fixed-size, 4.3 KB each);
    # in a training run these are the completed games' records selfplay_drain hands out
    if world > 1:
        comms[0].gather_replay(aug_recs)                    # first use: untimed
        t0 = time.perf_counter()
        all_recs = comms[0].gather_replay(aug_recs)
        comm_stats["gather_ms"] = 1e3 * red.max(time.perf_counter() - t0)
        comm_stats["gather_records"] = len(all_recs)
        comm_stats["gather_bytes"] = len(all_recs) * C.sizeof(tb.ReplayRecord)
        comm_stats["gather_what"] = ("selfplay_gather_replay: H2D of this rank's records, ncclAllGather of the counts and of "
                                     "the padded records, D2H of all ranks' records (host-timed, incl. the ctypes copies)")

    # ---------------- training step (next row N1): train_inner + all-reduce + Adam on those examples -----------------
    train = train_rate(engines[0], aug_tensors, world, comms[0], red, pk, aug_recs) if aug_tensors is not None else None
    del aug_tensors
    some_states = host_states[0][:256]

    # ---------------- the other north-star configs; the big engines' node pools are released first ----------------
    for r in range(1, E):
        if comms[r]:
            comms[r].close()
        engines[r].close()
    torch.cuda.empty_cache()
    movegen["perft5x5_d4"] = perft5_rate(local) if rank == 0 else None
    stress8 = stress8_rates(local, world, rank, red, pk)
    movegen["playouts8x8"] = stress8["playouts"]
    if comms[0]:
        comm_stats["bytes_moved_rank0_replica0"] = comms[0].bytes_moved()
    # engine 0 (and its communicator, which the reducer uses) stays alive until the sharded sub-benchmarks are done
    selfplay5 = selfplay5_rate(local, world, rank, red, max(1, min(args.steps, 3)), 2, W.random_weights(5, seed=0))

    # ---------------- small-batch regime of the analysis / playtak / pit callers (rank 0) ---------------
    if comms[0]:
        comms[0].close()
    engines[0].close()
    interactive = interactive_rates(local, blob, some_states) if rank == 0 else None
    late = late_game_rate(local, blob, max(1, min(args.steps, 3))) if rank == 0 else None

    if rank == 0 and world == 1:
        # the CPU path beside it: the oracle's literal restatement of perft.rs:3-18 (-O3 -march=native), single-threaded as
        # the reference's test is, and one subtree per host thread
        import oracle
        native = oracle.use_native_build()
        og = oracle.Game(6, 0)
        t0 = time.perf_counter()
        n4 = og.perft(4)
        t1 = time.perf_counter()
        threads = os.cpu_count() or 1
        n5 = og.perft(5, threads)
        t2 = time.perf_counter()
        movegen["cpu_baseline"] = {"kind": "port", "single_thread_mnodes_s": n4 / (t1 - t0) / 1e6,
                                   "all_threads_mnodes_s": n5 / (t2 - t1) / 1e6, "cores": threads,
                                   "march_native": bool(native),
                                   "sample": "6x6 perft(4) on one thread, perft(5) over all host threads; counts "
                                             f"{n4} / {n5}", "exact": n4 == 13_586_048 and n5 == PERFT6_D5}
    line = None
    if rank == 0:
        # the CPU baseline is timed beside the GPU arm at N=1 only (at N>1 the other ranks' host threads share the cores)
        cpu = cpu_selfplay_sample(max(8, min(R, 24)), R) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "6x6 batched self-play, 800 rollouts/move, random-init Net6 (configs[2])",
                       "games_per_gpu": G, "replicas_per_gpu": E, "rollouts": R,
                       "noise": "dirichlet(0.2) x0.3 below ply 80",
                       "pick": "visit-weighted sample below ply 40, argmax after", "instant_win": True,
                       "l2": "inputs larger than L2: every rollout step walks a node pool of tens of GB (12.6 MB per "
                             "game) and evaluates other leaves; the conv tower keeps a tile group's activations in L2 "
                             "by design and writes 36 KB of logits per leaf to HBM",
                       "parallelism": f"games sharded over {world} rank(s) x {E} engine replica(s), no collective in "
                                      "the rollout loop; weights / replay / gradients over NCCL through the C ABI"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "what": "host game states -> tak_games_upload -> mcts_rollouts(800) -> mcts_children_batch + "
                            "mcts_pick_move -> host"},
            "gpu_launches": total_launches,
            "roofline": roofline,
            "movegen": movegen,
            "selfplay5x5": selfplay5,
            "mcts8x8_dummy": stress8["mcts_dummy"],
            "comm": comm_stats or None,
            "augment": augment,
            "train": train,
            "interactive": interactive,
            "late_game": late,
            "cpu_baseline": {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                             "sample": cpu["sample"]} if cpu else None,
            "clocks": clocks,
            "extra": {"wall_ms_per_step": wall_max / args.steps, "wall_moves_per_s": total_plies / (wall_max / 1e3),
                      "evals_per_step": evals / max(1, args.steps), "launches_per_ply_per_replica": launches / max(1, args.steps) / E,
                      "games_completed": games_done, "replay_records_bytes": replay_bytes,
                      "net_evals_per_s": evals / (dev_ms / 1e3) if dev_ms else None},
        }
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=0,
                    help="concurrent games per GPU (default per replica: 148 SMs x 4 conv tiles x 7 boards = 4144)")
    ap.add_argument("--replicas", type=int, default=2, help="independent engine replicas per GPU (host thread each)")
    ap.add_argument("--rollouts", type=int, default=800)
    ap.add_argument("--nodes-per-game", type=int, default=1 << 18)
    args = ap.parse_args()
    # Native libraries may write to fd 1 (NCCL prints its version banner there when NCCL_DEBUG is set): everything but the
    # ONE JSON line goes to stderr.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
