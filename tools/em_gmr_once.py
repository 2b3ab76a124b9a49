"""One EM pass (K = 64, d = 7) and one GMR call (K = 8, O = 2) at reduced sizes, for ncu captures of k_gmm_em / k_gmr."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import perf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
print(perf.em(n))
print(perf.gmr(2 * n))
