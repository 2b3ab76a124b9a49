"""GMR kernel on the 2^26-point time grid and on random overlapping queries (perf.gmr) + parity tests."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import perf
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
r = perf.gmr(M)
print('grid_cfg1', r['grid_cfg1']); print('random_overlapping', r['random_overlapping'])
