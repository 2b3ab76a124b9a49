"""One GMR call on a 2^24-point time grid through the cfg-1 mixture (for ncu captures of k_gmr)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import perf
print(perf.gmr(1 << 24))
