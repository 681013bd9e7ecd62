"""Small driver for timing / ncu: perft from the opening through the C ABI.
  python tools/probe_perft.py [depth=5] [n=6] [reps=5]
Prints the count, the event-timed profile of every repetition (tak_perft_profile) and the derived rates:
materialised states/s over the expansion kernels, and SURVEY 8(d)'s S + S/b + 2 bytes per materialised state."""
import json
import sys
sys.path.insert(0, '.')
import tak_b200 as tb
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
S = {3: 160, 4: 192, 5: 288, 6: 384, 7: 896, 8: 1152}[n]
eng = tb.Engine(n, 64, nodes_per_game=64)
eng.reset(0, 1, 0)
root = eng.download([0])[0]
eng.perft(root, depth)       # allocates the frontier arenas
for _ in range(reps):
    nodes = eng.perft(root, depth)
    p = eng.perft_profile()
    b = p["materialised"] / max(1, p["materialised"] - p["top_children"]) if p["materialised"] else 0
    p["nodes"] = nodes
    p["counted_gnodes_s"] = nodes / (p["ms"] * 1e6)
    p["materialised_gstates_s_whole_call"] = p["materialised"] / (p["ms"] * 1e6)
    p["top_expand_gstates_s"] = p["top_children"] / (p["top_ms"] * 1e6) if p["top_ms"] else None
    p["top_expand_gbs_8d"] = p["top_children"] * (S + S / max(b, 1) + 2) / (p["top_ms"] * 1e6) if p["top_ms"] else None
    print(json.dumps(p))
eng.close()
