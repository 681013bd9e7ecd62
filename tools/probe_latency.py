"""Where the time of a small Network::policy_eval goes: device time of the forward (CUDA events, net_forward_profile) against
the wall time of the host-buffer call, for b = 1, 6, 32, 256."""
import sys
import time
sys.path.insert(0, '.')
import numpy as np  # noqa: E402
import tak_b200 as tb  # noqa: E402
from tak_b200 import weights as W  # noqa: E402

eng = tb.Engine(6, 256, nodes_per_game=64, max_batch=256)
eng.net_create(6)
eng.net_load_weights(W.random_weights(6, seed=0))
eng.reset(0, 256, 4)
states = eng.download(np.arange(256, dtype=np.int32))
for b in (1, 6, 32, 256):
    prof = eng.net_forward_profile(0, b, 20)
    st = (tb.TakState * b)(*states[:b])
    res = {}
    for name, buf in (("pinned", (eng.pinned_array((b, eng.policy_size)), eng.pinned_array((b,)))), ("pageable", None)):
        eng.policy_eval(st, out=buf)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            eng.policy_eval(st, out=buf)
            ts.append(time.perf_counter() - t0)
        res[name] = 1e3 * float(np.median(ts))
    print({"b": b, "device_forward_ms": prof["ms_forward"], "device_tower_ms": prof["ms_conv"], "policy_eval_wall_ms": res})
eng.close()
