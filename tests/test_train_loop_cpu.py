"""Host logic of the training loop (tak_b200/train_loop.py) without a GPU: the `Network::train` schedule
(network.rs:47-56: shuffled references, chunks_exact, a step every CHUNKS_IN_STEP chunks, trailing gradients dropped) and its
data-parallel dealing (chunk i -> rank i % world, all ranks step together) on a recording stand-in for the engine."""
import numpy as np

from tak_b200 import train_loop as TL


class FakeEngine:
    def __init__(self):
        self.log = []

    def train_begin(self, boards):
        self.log.append(("begin", boards))

    def examples_to_tensors(self, recs, on_device=False):
        assert on_device
        return list(recs), None, None

    def train_chunk(self, inputs, pi, z):
        self.log.append(("chunk", tuple(inputs)))
        return 1.0, 0.5

    def train_grad_tensor(self):
        return "grad"

    def train_step(self, lr, wd):
        self.log.append(("step", lr, wd))

    def train_get(self, what):
        return np.zeros(3, dtype=np.float32)

    def train_end(self):
        self.log.append(("end",))


def run(rank, world, n_examples=23, chunk=4, cis=2):
    eng, reduced = FakeEngine(), []
    TL.train_network(eng, list(range(n_examples)), np.random.default_rng(7), chunk_size=chunk, chunks_in_step=cis,
                     allreduce=lambda e: reduced.append(e.train_grad_tensor()), log=lambda *_: None, rank=rank, world=world)
    return eng.log, reduced


def test_single_process_schedule_follows_network_train():
    log, reduced = run(0, 1)
    chunks = [e[1] for e in log if e[0] == "chunk"]
    assert log[0] == ("begin", 32) and log[-1] == ("end",)
    assert len(chunks) == 23 // 4 == 5 and all(len(c) == 4 for c in chunks)        # chunks_exact: 3 examples are left out
    seen = [x for c in chunks for x in c]
    assert len(set(seen)) == 20 and seen != sorted(seen)                          # shuffled, no example twice
    kinds = [e[0] for e in log]
    assert kinds == ["begin", "chunk", "chunk", "step", "chunk", "chunk", "step", "chunk", "end"]   # 5th chunk: no step
    assert all(e[1:] == (TL.LEARNING_RATE, TL.WEIGHT_DECAY) for e in log if e[0] == "step")
    assert reduced == ["grad", "grad"]                                             # all-reduce right before each step


def test_data_parallel_dealing_covers_every_chunk_once():
    single = [e[1] for e in run(0, 1)[0] if e[0] == "chunk"]
    per_rank = [run(r, 2)[0] for r in range(2)]
    dealt = [[e[1] for e in log if e[0] == "chunk"] for log in per_rank]
    assert dealt[0] == single[0::2] and dealt[1] == single[1::2]                   # same shuffle, chunk i -> rank i % 2
    for log in per_rank:                                                           # every rank steps at the same points
        assert [e[0] for e in log].count("step") == 2
    # with world == chunks_in_step each rank contributes exactly one chunk per step
    k0 = [e[0] for e in per_rank[0]]
    assert k0 == ["begin", "chunk", "step", "chunk", "step", "chunk", "end"]
