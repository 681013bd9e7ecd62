"""Small driver for ncu: one 6x6 perft from the opening through the C ABI (depth from argv, default 5)."""
import sys
sys.path.insert(0, '.')
import tak_b200 as tb
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
eng = tb.Engine(n, 64, nodes_per_game=64)
eng.reset(0, 1, 0)
root = eng.download([0])[0]
eng.perft(root, 3)
print(eng.perft(root, depth), eng.perft_stats())
eng.close()
