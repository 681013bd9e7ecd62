"""The `train` binary's loop on the device engines (train/src/main.rs:82-123, alpha-tak/src/model/network.rs:37-57):

    loop { if examples: new = copy(network); new.train(examples); if pit(new, network).win_rate() > 0.55: network = new
           examples += self_play_parallel(network) }

`train_network` is `Network::train` (fresh Adam, shuffled references, chunks_exact(CHUNK_SIZE), a step every
CHUNKS_IN_STEP chunks -- gradients of a trailing partial group are dropped, as in the reference); `training_iteration` is
one turn of `training_loop`.  Everything heavy runs behind the C ABI (self-play, augmentation, train_inner, Adam, pit);
this file is the cold control flow.  With `save_dir` the reference's files are written too: `_models/<time>.model` (tch
VarStore archive, weights.save_tch_model) when a candidate is accepted and `_examples/<time>.data` (one Example per line,
example.rs:81-100) after every self-play round.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from ._lib import ReplayRecord
from .engine import Engine, example_format
from .pit import PitResult, pit

CHUNK_SIZE = 500            # network.rs:19
CHUNKS_IN_STEP = 20         # network.rs:20
LEARNING_RATE = 1e-4        # network.rs:14
WEIGHT_DECAY = 1e-4         # network.rs:15
MAX_EXAMPLES = 400_000      # train/src/main.rs:25
WIN_RATE_THRESHOLD = 0.55   # train/src/main.rs:27
# self_play_parallel's constants (train/src/self_play.rs:12-18): what `training_iteration` / `collect_self_play` use
# unless the caller overrides them.  (Engine.selfplay_begin's own defaults -- noise off -- are the PARITY configuration.)
SELF_PLAY_DEFAULTS = dict(rollouts=10_000, noise_ply=80, noise_alpha=0.2, noise_ratio=0.3, exploit_ply=40, instant_win=1,
                          half_komi=4)


def train_network(eng: Engine, examples: Sequence[ReplayRecord], rng: np.random.Generator,
                  chunk_size: int = CHUNK_SIZE, chunks_in_step: int = CHUNKS_IN_STEP, lr: float = LEARNING_RATE,
                  weight_decay: float = WEIGHT_DECAY, allreduce: Optional[Callable] = None,
                  log: Callable = print, rank: int = 0, world: int = 1) -> np.ndarray:
    """`Network::train(&mut self, examples)` starting from the weights last loaded into `eng`; returns the new blob
    (weights incl. BatchNorm running statistics).

    Data-parallel (world > 1): every rank holds the same examples and the same `rng` state, so all ranks draw the same
    shuffle; chunk i is trained by rank i % world, and before every step `allreduce(eng)` (parallel.engine_allreduce:
    net_train_allreduce on the engine's stream, or the synchronised torch.distributed fallback) sums the ranks'
    accumulators -- the step is then the reference's single-process step over the same `chunks_in_step` chunks.
    (BatchNorm running statistics follow each rank's own chunks.)"""
    log(f"starting training with {len(examples)} examples")
    eng.train_begin(8 * chunk_size)                       # Adam { wd, ..Default }.build(vs, lr): fresh moments
    order = rng.permutation(len(examples))                # refs.shuffle(&mut thread_rng())
    losses = []
    for i in range(len(examples) // chunk_size):          # chunks_exact(CHUNK_SIZE)
        if i % world == rank:
            recs = [examples[j] for j in order[i * chunk_size:(i + 1) * chunk_size]]
            inputs, pi, z = eng.examples_to_tensors(recs, on_device=True)    # flat_map(|ex| ex.to_tensors())
            lp, lz = eng.train_chunk(inputs, pi, z)
            losses.append((lp, lz))
            log(f"p={lp:.4f}\t z={lz:.4f}")
        if (i + 1) % chunks_in_step == 0:
            if allreduce is not None:
                allreduce(eng)                # sums the ranks' gradient accumulators; ordered before the step
            log("making step!")
            eng.train_step(lr, weight_decay)
    blob = eng.train_get(0)
    eng.train_end()
    train_network.last_losses = losses
    return blob


def collect_self_play(eng: Engine, min_examples: int, max_steps: int = 10_000, **selfplay_cfg) -> List[ReplayRecord]:
    """`self_play_parallel(&network)` until at least `min_examples` completed-game records exist (the reference plays a
    fixed number of games; the device loop keeps every slot busy and hands back the records of finished games)."""
    eng.selfplay_begin(**{**SELF_PLAY_DEFAULTS, **selfplay_cfg})
    out: List[ReplayRecord] = []
    for _ in range(max_steps):
        eng.selfplay_step(1)
        out += eng.selfplay_drain()
        if len(out) >= min_examples:
            break
    return out


def save_examples(path: str, examples: Sequence[ReplayRecord]) -> None:
    with open(path, "w") as f:
        for ex in examples:
            f.write(example_format(ex) + "\n")


def training_iteration(current: Engine, candidate: Engine, blob: np.ndarray, examples: List[ReplayRecord],
                       rng: np.random.Generator, pit_games: int = 128, pit_rollouts: int = 50, pit_batch: int = 16,
                       min_new_examples: int = 1000, train_kw: Optional[dict] = None,
                       selfplay_kw: Optional[dict] = None, log: Callable = print, save_dir: Optional[str] = None):
    """One turn of `training_loop` (train/src/main.rs:82-123).  `current` holds the accepted network (weights `blob`),
    `candidate` is a second engine for the copy being trained and pitted.  Returns (blob, examples, PitResult | None)."""
    result: Optional[PitResult] = None
    if examples:
        candidate.net_load_weights(blob)                  # copy(&network)
        new_blob = train_network(candidate, examples, rng, log=log, **(train_kw or {}))
        candidate.net_load_weights(new_blob)
        log("pitting two networks against each other")
        result = pit(candidate, current, games=pit_games, batch=pit_batch, rollouts=pit_rollouts,
                     seed=int(rng.integers(1 << 31)))
        log(repr(result))
        if result.win_rate() > WIN_RATE_THRESHOLD:
            blob = new_blob
            current.net_load_weights(blob)
            log("saving model")
            if save_dir:                                  # network.save("_models/<unix time>.model") (main.rs:104)
                import os
                import time

                from . import weights as W
                os.makedirs(os.path.join(save_dir, "_models"), exist_ok=True)
                W.save_tch_model(blob, current.n, os.path.join(save_dir, "_models", f"{int(time.time())}.model"))
        if len(examples) > MAX_EXAMPLES:
            examples = examples[-MAX_EXAMPLES:]
    log("starting self-play")
    new_examples = collect_self_play(current, min_new_examples, **(selfplay_kw or {}))
    if save_dir:                                          # "_examples/<unix time>.data" (self_play.rs:98,253-255)
        import os
        import time
        os.makedirs(os.path.join(save_dir, "_examples"), exist_ok=True)
        save_examples(os.path.join(save_dir, "_examples", f"{int(time.time())}.data"), new_examples)
    examples = list(examples) + new_examples
    return blob, examples, result


def distributed_iteration(current: Engine, candidate: Engine, blob: np.ndarray, examples: List[ReplayRecord], seed: int,
                          device, pit_games: int = 128, pit_rollouts: int = 50, pit_batch: int = 16,
                          min_new_examples: int = 1000, train_kw: Optional[dict] = None,
                          selfplay_kw: Optional[dict] = None, log: Callable = print, comm=None):
    """One turn of `training_loop` on N GPUs (one process per GPU, torch.distributed already initialised):
      * training is data-parallel (chunks dealt round-robin, NCCL all-reduce of the gradient blob before each step),
      * the pit games are split over the ranks and the win / loss / draw counts summed,
      * self-play is sharded (every rank plays its own games with the SAME accepted network) and the replay records are
        all-gathered, so every rank enters the next turn with the same examples and the same weights.
    Returns (blob, examples, PitResult | None); identical on every rank."""
    import torch
    import torch.distributed as dist

    from . import parallel as par
    rank, world = dist.get_rank(), dist.get_world_size()
    result: Optional[PitResult] = None
    if examples:
        candidate.net_load_weights(blob)
        rng = np.random.default_rng(seed)                 # same shuffle on every rank
        # `comm`: a tak_b200.comm.Comm on `candidate` (NCCL on its stream); without one, torch.distributed + a sync
        new_blob = train_network(candidate, examples, rng, log=log,
                                 allreduce=lambda e: par.engine_allreduce(e, comm), rank=rank, world=world,
                                 **(train_kw or {}))
        t = torch.from_numpy(new_blob).to(device)         # running statistics: mean over ranks; weights already agree
        dist.all_reduce(t)
        new_blob = (t / world).cpu().numpy()
        candidate.net_load_weights(new_blob)
        mine = pit(candidate, current, games=max(1, pit_games // world), batch=pit_batch, rollouts=pit_rollouts,
                   seed=seed * 1000 + rank)
        counts = torch.tensor([mine.wins, mine.losses, mine.draws], dtype=torch.int64, device=device)
        dist.all_reduce(counts)
        result = PitResult(*[int(c) for c in counts.tolist()])
        log(repr(result))
        if result.win_rate() > WIN_RATE_THRESHOLD:
            blob = new_blob
            current.net_load_weights(blob)
        if len(examples) > MAX_EXAMPLES:
            examples = examples[-MAX_EXAMPLES:]
    sp = dict(selfplay_kw or {})
    sp["game_id_base"] = par.game_id_base(rank, current.max_games)
    sp["seed"] = sp.get("seed", 0x7A4B) + seed
    mine = collect_self_play(current, max(1, min_new_examples // world), **sp)
    examples = list(examples) + par.gather_replay(mine, ReplayRecord, device)
    return blob, examples, result
