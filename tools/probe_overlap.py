"""The headline self-play step in isolation: `replicas` engines x `games` games, 800 rollouts, bench.py's settings;
prints moves/s (games searched per second over the slower replica's wall time).  For A/B runs of the search-step /
tower overlap switches (TAK_STEP_WARPS, TAK_STEP_STREAM, TAK_TOWER_PRIO, TAK_STEP_TIMING).
Usage: probe_overlap.py [replicas=2] [games_per_replica=4144] [steps=3] [warmup=2]"""
import sys
import threading
import time

sys.path.insert(0, '.')
import tak_b200 as tb  # noqa: E402
from tak_b200 import weights as W  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Gr = int(sys.argv[2]) if len(sys.argv) > 2 else 4144
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
warmup = int(sys.argv[4]) if len(sys.argv) > 4 else 2
blob = W.random_weights(6, seed=0)
engines = []
for r in range(E):
    eng = tb.Engine(6, Gr, nodes_per_game=1 << 18, max_batch=Gr)
    eng.net_create(6)
    eng.net_load_weights(blob)
    eng.selfplay_begin(rollouts=800, half_komi=4, instant_win=1, exploit_ply=40, noise_ply=80, noise_alpha=0.2,
                       noise_ratio=0.3, seed=0x7A4B, game_id_base=r * Gr)
    engines.append(eng)


def run(fns):
    ts = [threading.Thread(target=f) for f in fns]
    [t.start() for t in ts]
    [t.join() for t in ts]


def loop(eng, k, out):
    ms = 0.0
    for _ in range(k):
        st = eng.selfplay_step(1)
        ms += st.device_ms
        eng.selfplay_drain()
    out.append(ms)


run([lambda e=e: loop(e, warmup, []) for e in engines])
outs = [[] for _ in engines]
t0 = time.perf_counter()
run([lambda e=e, o=o: loop(e, steps, o) for e, o in zip(engines, outs)])
wall = time.perf_counter() - t0
dev_ms = max(o[0] for o in outs)
print({"replicas": E, "games": E * Gr, "steps": steps, "moves_per_s_device": E * Gr * steps / (dev_ms * 1e-3),
       "moves_per_s_wall": E * Gr * steps / wall, "ms_per_step": dev_ms / steps})
for e in engines:
    e.close()
