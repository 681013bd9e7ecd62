"""2-GPU test of the NCCL entry points of the C ABI (tak_comm_init, net_broadcast_weights, selfplay_gather_replay,
net_train_allreduce, tak_comm_sum_u64 / max_f64) driven through ctypes by two plain processes -- no torch.distributed
anywhere; the unique id travels through a multiprocessing pipe.  Needs two visible GPUs (`gpurun --gpus 2`); on a
single-GPU box it is skipped (NCCL refuses two ranks on one device)."""
import multiprocessing as mp

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, uid_pipe, out_q):
    try:
        import tak_b200 as tb
        from tak_b200 import comm as tc
        from tak_b200 import weights as W
        uid = tc.unique_id() if rank == 0 else None
        if rank == 0:
            for p in uid_pipe:
                p.send(uid)
        else:
            uid = uid_pipe.recv()
        G = 8
        eng = tb.Engine(6, G, device=rank, nodes_per_game=1 << 12, max_batch=64)
        eng.net_create(6)
        cm = tc.Comm(eng, uid, rank, world)
        # 1. weights: only rank 0 has them
        blob = W.random_weights(6, seed=31) if rank == 0 else None
        cm.broadcast_weights(blob, root=0)
        st = tb.state_init(6, 4)
        pol, val = eng.policy_eval([st])
        # 2. self-play on disjoint game ids, replay gathered on every rank
        eng.selfplay_begin(rollouts=8, half_komi=4, instant_win=1, exploit_ply=0, noise_ply=0, seed=5, max_plies=6,
                           game_id_base=rank * G)
        mine = []
        for _ in range(8):
            eng.selfplay_step(1)
            mine += eng.selfplay_drain(4096)
        allrec = cm.gather_replay(mine)
        # 3. data-parallel gradient sum: each rank trains one chunk of its own examples
        eng.net_load_weights(W.random_weights(6, seed=31))
        eng.train_begin(64)
        inputs, pi, z = eng.examples_to_tensors(mine[:8], on_device=True)
        eng.train_chunk(inputs, pi, z)
        g_local = eng.train_get(1)
        cm.allreduce_gradients()
        g_sum = eng.train_get(1)
        eng.train_step(1e-4, 1e-4)
        w_after = eng.train_get(0)
        eng.train_end()
        # 4. the same through train_loop.train_network (Network::train, data-parallel): chunk i goes to rank i % world and
        #    the hook sums the accumulators on the engine stream right before each Adam step -- no host synchronisation
        #    between the all-reduce and the step (the race the torch.distributed route had)
        import numpy as np
        from tak_b200 import train_loop as TL
        eng.net_load_weights(W.random_weights(6, seed=31))
        tn_blob = TL.train_network(eng, allrec[:32], np.random.default_rng(3), chunk_size=8, chunks_in_step=2, lr=1e-3,
                                   allreduce=lambda e: cm.allreduce_gradients(), log=lambda *_: None, rank=rank,
                                   world=world)
        total = cm.sum_u64([len(mine), 7])
        mx = cm.max_f64([float(rank), 1.5])
        moved = cm.bytes_moved()
        cm.close()
        eng.close()
        out_q.put((rank, {"pol": pol[0][:64].copy(), "val": float(val[0]), "n_mine": len(mine),
                          "ids": [(r.game_id, r.game_serial, int(r.state.ply)) for r in allrec],
                          "g_local": g_local, "g_sum": g_sum, "w_after": w_after, "tn_blob": tn_blob, "total": total, "mx": mx,
                          "moved": moved}))
    except BaseException as ex:  # noqa: BLE001
        import traceback
        out_q.put((rank, {"error": f"{ex}\n{traceback.format_exc()}"}))


def test_nccl_entry_points_two_ranks_no_torch_distributed():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    pipes = [ctx.Pipe(duplex=False) for _ in range(world - 1)]
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(0, world, [w for _, w in pipes], q))]
    procs += [ctx.Process(target=_worker, args=(r, world, pipes[r - 1][0], q)) for r in range(1, world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert "error" not in res[r], res[r].get("error")
    a, b = res[0], res[1]
    # broadcast: both ranks evaluate the opening position to the same bits with the network only rank 0 had
    assert np.array_equal(a["pol"].view(np.uint32), b["pol"].view(np.uint32)) and a["val"] == b["val"]
    assert a["pol"].std() > 0
    # gather: every rank holds all ranks' records, rank 0's first; ids are disjoint by construction
    assert a["ids"] == b["ids"] and len(a["ids"]) == a["n_mine"] + b["n_mine"] > 0
    assert all(gid < 8 for gid, _, _ in a["ids"][:a["n_mine"]]) and all(gid >= 8 for gid, _, _ in a["ids"][a["n_mine"]:])
    # all-reduce: the summed blob is the sum of the two local blobs, identical on both ranks, and so is the Adam step
    # (trainable tensors: the BatchNorm running statistics in the blob follow each rank's own chunk)
    from tak_b200 import weights as W
    mask = np.concatenate([np.full(int(np.prod(s)), "running_" not in n) for n, s in W.spec(6)])
    assert np.array_equal(a["g_sum"], b["g_sum"]) and np.array_equal(a["w_after"][mask], b["w_after"][mask])
    assert not np.array_equal(a["w_after"][~mask], b["w_after"][~mask])
    want = a["g_local"].astype(np.float64) + b["g_local"].astype(np.float64)
    assert np.abs(a["g_sum"] - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    assert np.abs(a["g_local"]).max() > 0 and not np.array_equal(a["g_local"], b["g_local"])
    # train_network on two ranks: identical trainable weights on both (BatchNorm running statistics follow each rank's own
    # chunks), and they moved away from the start
    assert np.array_equal(a["tn_blob"][mask], b["tn_blob"][mask])
    assert not np.array_equal(a["tn_blob"][mask], W.random_weights(6, seed=31)[mask])
    assert a["total"] == b["total"] == [a["n_mine"] + b["n_mine"], 14]
    assert a["mx"] == b["mx"] == [1.0, 1.5]
    assert a["moved"] > 20_000_000            # weight blob + gradient blob at least
