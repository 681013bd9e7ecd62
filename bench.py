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


def movegen_mnodes(eng, world, rank, dev, pk):
    """The metric's second half: movegen + play + result throughput as perft(5) of the 6x6 opening position through the
    C ABI (tak_perft), breadth-first on the device.  With N ranks the 1 260 positions two plies below the root are dealt
    round-robin (tak_perft_multi expands a rank's share as one frontier) and the counts are summed with one all-reduce
    (SURVEY.md section 8e)."""
    from tak_b200 import parallel as par

    depth = 5
    eng.reset(0, 1, 0)                       # slot 0 of this rank's engine (self-play is over) holds the opening
    root = eng.download([0])[0]
    ms, nodes, mat = 0.0, 0, 0
    eng.perft(root, depth)                   # warm-up at full depth: the frontier arenas are allocated here
    if world == 1:
        nodes = eng.perft(root, depth)
        st = eng.perft_stats()
        ms, mat = st["ms"], st["materialised"]
    else:
        # every rank builds the depth-2 frontier (1 260 positions, host-driven, untimed), takes every world-th position
        # and expands its share in ONE breadth-first perft of the remaining depth
        front, ended = eng.frontier(root, 2)
        mine = front[rank::world]
        eng.perft_multi(mine, depth - 2)     # warm-up: arenas sized for this share
        nodes = eng.perft_multi(mine, depth - 2) + (ended if rank == 0 else 0)
        st = eng.perft_stats()
        ms, mat = st["ms"], st["materialised"]
    total = int(par.sum_over_ranks(float(nodes), dev))
    t_max = par.max_over_ranks(ms, dev)
    mat_total = par.sum_over_ranks(float(mat), dev)
    S = 384                                   # packed 6x6 state bytes
    b = total / mat_total if mat_total else 0.0   # ~ mean branching of the counted level (92 from the opening)
    # HBM bytes the breadth-first expansion must move: every materialised node is written once (S + 2 B move) and read
    # once by the next level's count and once by its expand; the 1.25e9 leaves are only counted on chip
    algo_bytes = mat_total * (3 * S + 2)
    gbs = algo_bytes / (t_max * 1e-3) / 1e9 if t_max else 0.0
    return {"value": total / (t_max * 1e-3) / 1e6 if t_max else None, "unit": "Mnodes/s",
            "workload": "6x6 perft depth 5 from the opening via tak_perft (movegen + play + result, bit-exact count)"
                        + ("" if world == 1 else f"; the 1260 positions two plies down are dealt round-robin over {world} ranks "
                           "(that shallow frontier is built on the host, untimed), each rank's share is one tak_perft_multi"),
            "nodes": total, "exact": total == PERFT6_D5, "ms": t_max, "materialised_states": int(mat_total),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": gbs / pk["hbm"] if pk["hbm"] else None,
                         "note": "13.7 M states are materialised (depth <= 4: written once, read by the count and the "
                                 "expand kernels); the 1.25e9 leaves are only counted, one thread per depth-4 parent from "
                                 "the 96-byte tail of its record (closed-form move counts)",
                         "mean_branching_last_level": b}}


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


def interactive_rates(local, blob_ptr, elems, states):
    """The `analysis` / `playtak` / `pit` regime (one game, small batches): latency of Network::policy_eval through the
    host-buffer ABI (alpha-tak/src/model/network.rs:34) and the rollouts/s of a single `Player` with batch 32
    (analysis/src/main.rs prints the same figure as nps)."""
    import tak_b200 as tb

    # a search of one game for a second grows a tree of millions of nodes: its own engine with a deep node pool
    eng = tb.Engine(6, 2, device=local, nodes_per_game=1 << 23, max_batch=256)
    eng.net_create(6)
    eng.net_load_weights_device(blob_ptr, elems)
    out = {"policy_eval_ms": {}, "what": "net_policy_eval(host states -> full softmax [b, 9036] + value on the host), "
                                         "median of 20 calls; Player(batch 32).rollout() on one game for 1 s"}
    for b in (1, 32, 256):
        st = states[:b]
        eng.policy_eval(st)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            eng.policy_eval(st)
            ts.append(time.perf_counter() - t0)
        out["policy_eval_ms"][str(b)] = 1e3 * float(np.median(ts))
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
    return out


def train_rate(eng, tensors, world, dist, dev, pk, host_recs=None):
    """Network::train_inner + Adam (next row N1, network.rs:37-97) on the augmented examples of this run: chunks of
    500 examples x 8 symmetries = 4000 positions (CHUNK_SIZE, network.rs:19), inputs resident in HBM; with N ranks every
    rank trains its own chunks and the fp32 gradient blob is all-reduced over NCCL before the Adam step."""
    import torch

    from tak_b200 import parallel as par_mod
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
        dt = par_max((time.perf_counter() - t0) / reps, dev)
        e2e = {"value": world * B / dt, "unit": "positions/s", "ms_per_chunk": 1e3 * dt,
               "h2d_bytes_per_step": len(recs) * C.sizeof(type(recs[0])), "d2h_bytes_per_step": 8,
               "what": "host replay records -> examples_to_tensors(on_device) -> net_train_chunk -> (loss_p, loss_z)"}
    g = eng.train_grad_tensor()
    if dist:                                  # untimed: NCCL builds its rings / buffers for this size on first use
        par_mod.allreduce_gradients(torch.zeros_like(g))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if dist:
        par_mod.allreduce_gradients(g)
        torch.cuda.synchronize()
    t_ar = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.train_step(1e-4, 1e-4)
    t_step = time.perf_counter() - t0
    ms_chunk = float(np.mean(ms))
    ms_max = par_max(ms_chunk, dev)
    flop = 3.0 * FLOP_PER_EVAL_NET6 * B          # forward + dgrad + wgrad
    out = {"value": world * B / (ms_max * 1e-3), "unit": "positions/s", "positions_per_chunk": B,
           "ms_per_chunk": ms_max, "allreduce_ms": 1e3 * t_ar, "adam_step_ms": 1e3 * t_step,
           "grad_bytes": int(W.blob_size(6)) * 4, "loss_p": loss[0], "loss_z": loss[1],
           "tflops": flop / (ms_max * 1e-3) / 1e12, "tensor_peak": pk["bf16_sustained"],
           "frac_of_tensor_peak": flop / (ms_max * 1e-3) / 1e12 / pk["bf16_sustained"], "e2e": e2e,
           "what": "train_inner on 4000 augmented positions (forward_training + loss + backward, CUDA events on the engine "
                   "stream), then NCCL all-reduce of the gradient blob and the Adam step"}
    eng.train_end()
    return out


def par_max(x, dev):
    from tak_b200 import parallel as par
    return par.max_over_ranks(x, dev)


def run_b200(args):
    import torch

    import tak_b200 as tb
    from tak_b200 import parallel as par
    from tak_b200 import weights as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    # Games never interact, so a GPU's games are split over `replicas` independent engines (own stream, search trees and
    # network replica) driven by one host thread each: while one replica's conv tower owns the tensor cores, the other's
    # MCTS / encode / head kernels run beside it on the same SMs.
    E = max(1, args.replicas)
    G = args.games if args.games else 5328 * E          # games per GPU; 5328 = 148 SMs x 6 tiles x 6 boards
    Gr = G // E
    G = Gr * E
    R, n = args.rollouts, 6
    pk = peaks()
    elems = W.blob_size(6)
    # weights: rank 0 draws them, NCCL broadcasts the fp32 blob over NVLink, every replica folds/packs its own copy
    blob_dev = par.broadcast_weights(W.random_weights(6, seed=0) if rank == 0 else None, elems, dev)
    torch.cuda.synchronize()
    engines = []
    for r in range(E):
        eng = tb.Engine(n, Gr, device=local, nodes_per_game=args.nodes_per_game, max_batch=Gr)
        eng.net_create(6)
        eng.net_load_weights_device(blob_dev.data_ptr(), elems)
        engines.append(eng)

    def barrier():
        torch.cuda.synchronize()
        for eng in engines:
            eng.sync()
        if dist:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        return par.max_over_ranks(x, dev)

    def sum_over_ranks(x: float) -> float:
        return par.sum_over_ranks(x, dev)

    # the conv tower timed ALONE on a cool GPU (3 launches, CUDA events) -> compared with the burst peak below
    engines[0].reset(0, Gr, 4)
    prof_alone = engines[0].net_forward_profile(0, Gr, 3)

    # ---------------- device-resident self-play: `value` ----------------
    for r, eng in enumerate(engines):
        eng.selfplay_begin(rollouts=R, half_komi=4, instant_win=1, exploit_ply=40, noise_ply=80, noise_alpha=0.2,
                           noise_ratio=0.3, seed=0x7A4B, game_id_base=par.game_id_base(rank * E + r, Gr))

    def warm(eng):
        for _ in range(args.warmup):
            eng.selfplay_step(1)
            eng.selfplay_drain(4 * Gr)

    in_threads([lambda eng=eng: warm(eng) for eng in engines])

    def timed(eng):
        """K searched plies of every game of this replica; device time = CUDA events on the replica's stream."""
        tot = {"ms": 0.0, "launches": 0, "evals": 0, "plies": 0, "done": 0, "recs": []}
        left = args.steps
        while left > 0:
            k = min(left, 4)                       # the replay ring holds 16 records per game between drains
            st = eng.selfplay_step(k)
            tot["ms"] += st.device_ms
            tot["launches"] += st.kernel_launches
            tot["evals"] += st.evals
            tot["plies"] += st.plies_played
            tot["done"] += st.games_completed
            left -= k
            if left > 0:
                tot["recs"] += eng.selfplay_drain(16 * Gr)
        return tot

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    res = in_threads([lambda eng=eng: timed(eng) for eng in engines])
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop()
    dev_ms = max(r["ms"] for r in res)                  # replicas run concurrently: the GPU's time is the longest span
    launches = sum(r["launches"] for r in res)
    evals = sum(r["evals"] for r in res)
    plies = sum(r["plies"] for r in res)
    games_done = sum(r["done"] for r in res)
    recs = sum((r["recs"] for r in res), [])
    for eng in engines:
        recs += eng.selfplay_drain(16 * Gr)
    # replay gather: fixed-size records, all-gathered over NCCL (outside the timed rollouts, as the trainer would)
    all_recs = par.gather_replay(recs, tb.ReplayRecord, dev)
    replay_bytes = len(all_recs) * C.sizeof(tb.ReplayRecord)
    t_max = max_over_ranks(max(dev_ms, 0.0))           # device time (CUDA events on the engine streams), max over ranks
    total_plies = sum_over_ranks(float(plies))
    value = total_plies / (t_max / 1e3)
    total_launches = int(sum_over_ranks(float(launches)))

    # ---------------- end to end through the host-buffer ABI: `e2e` ----------------
    ids = np.arange(Gr, dtype=np.int32)
    host_states = [eng.download(ids) for eng in engines]  # the positions self-play reached, now in host memory
    state_bytes = {5: 288, 6: 384}.get(n, 384)
    h2d = G * state_bytes + G * 4
    stride = 256
    d2h = G * stride * 6 + G * 4 + G * 2
    e2e_steps = max(1, min(args.steps, 2))

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
    e2e_t = max_over_ranks(time.perf_counter() - t0)
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
        # dram__bytes_read.sum + dram__bytes_write.sum of one tower launch over 5328 boards (ncu --set full,
        # profiles/r01_conv_tc3_ncu_full.txt), scaled to this launch's boards: logits + write-backs of the activations
        "traffic": 829_248_512 * Gr / 5328,
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

    # ---------------- movegen Mnodes/s: 6x6 perft(5) from the opening, root moves sharded over ranks ----------------
    movegen = movegen_mnodes(engines[0], world, rank, dev, pk)

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

    # ---------------- small-batch regime of the analysis / playtak / pit callers (rank 0's first replica) ---------------
    interactive = interactive_rates(local, blob_dev.data_ptr(), elems, host_states[0]) if rank == 0 else None

    # ---------------- training step (next row N1): train_inner + all-reduce + Adam on those examples -----------------
    train = train_rate(engines[0], aug_tensors, world, dist, dev, pk, aug_recs) if aug_tensors is not None else None

    if rank == 0 and world == 1:
        # the CPU path beside it: the oracle's literal restatement of perft.rs:3-18 (-O3 -march=native), single-threaded as
        # the reference's test is, and one subtree per host thread
        import oracle
        og = oracle.Game(6, 0)
        t0 = time.perf_counter()
        n4 = og.perft(4)
        t1 = time.perf_counter()
        threads = os.cpu_count() or 1
        n5 = og.perft(5, threads)
        t2 = time.perf_counter()
        movegen["cpu_baseline"] = {"kind": "port", "single_thread_mnodes_s": n4 / (t1 - t0) / 1e6,
                                   "all_threads_mnodes_s": n5 / (t2 - t1) / 1e6, "cores": threads,
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
                                      "the rollout loop"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "host game states -> tak_games_upload -> mcts_rollouts(800) -> mcts_children_batch + "
                            "mcts_pick_move -> host"},
            "gpu_launches": total_launches,
            "roofline": roofline,
            "movegen": movegen,
            "augment": augment,
            "train": train,
            "interactive": interactive,
            "cpu_baseline": {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                             "sample": cpu["sample"]} if cpu else None,
            "clocks": clocks,
            "extra": {"wall_ms_per_step": wall_ms / args.steps, "evals_per_step": evals / max(1, args.steps),
                      "games_completed": games_done, "replay_records_gathered_bytes": replay_bytes,
                      "net_evals_per_s": evals / (dev_ms / 1e3) if dev_ms else None},
        }
    for eng in engines:
        eng.close()
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
                    help="concurrent games per GPU (default 5328 per replica = 148 SMs x 6 conv tiles x 6 boards)")
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
