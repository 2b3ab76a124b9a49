"""Device time of the two kernels of a Lloyd iteration (assignment, hard-label statistics) and of one k-means++ centre
(sampling, candidate distances, greedy choice) at cfg 5's size (10 M x 7, K = 64)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib, perf
from kmp_demos_b200.mixture import stats_len
lib = _lib.load()
n, K, d = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 64, 7
dev = torch.device('cuda')
X, lab = perf.cfg5_on_device(n, K, d, 0, dev)
st = _lib.stream()
cent = X[torch.randint(0, n, (K,), device=dev)].contiguous()
labels = torch.empty(n, dtype=torch.int32, device=dev); nch = torch.zeros(1, dtype=torch.int32, device=dev)
ws_bytes = max(int(lib.kmpx_gmm_workspace_bytes(n, d, K)), 148 * 4 * 16 * 8, int(lib.kmpx_kpp_sample_workspace_bytes(n)))
ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.float64, device=dev)
stats = torch.zeros(stats_len(K, d), dtype=torch.float64, device=dev)
t = perf.timeit(lambda: _lib.call('kmpx_kmeans_assign', _lib.ptr(X), n, d, K, _lib.ptr(cent), _lib.ptr(lab), _lib.ptr(labels), _lib.ptr(nch), st))
print(f'kmeans_assign: {t:.3f} ms  ({3.0 * n * K * d / t * 1e-9:.2f} TFLOP/s at 3nKd)')
t = perf.timeit(lambda: _lib.call('kmpx_gmm_label_stats', _lib.ptr(X), n, d, K, None, _lib.ptr(labels), _lib.ptr(stats), _lib.ptr(ws), ws_bytes, st))
print(f'label_stats:   {t:.3f} ms')
x_sq = (X * X).sum(dim=1).contiguous()
closest = torch.rand(n, dtype=torch.float64, device=dev)
u = torch.rand(8, dtype=torch.float64, device=dev); pot = closest.sum().reshape(1)
cand = torch.zeros(8, dtype=torch.int64, device=dev); flags = torch.zeros(9, dtype=torch.int32, device=dev)
t = perf.timeit(lambda: _lib.call('kmpx_kpp_sample', _lib.ptr(closest), n, _lib.ptr(u), 6, _lib.ptr(pot), None, 0, n, 1, _lib.ptr(cand), _lib.ptr(flags), _lib.ptr(ws), ws_bytes, st))
print(f'kpp_sample:    {t:.3f} ms  flags {flags[:6].tolist()}')
rows = torch.zeros((8, d + 2), dtype=torch.float64, device=dev); rows[:6, :d] = X[cand[:6]]
dist = torch.empty((6, n), dtype=torch.float64, device=dev); potl = torch.zeros((6, 2), dtype=torch.float64, device=dev)
t = perf.timeit(lambda: _lib.call('kmpx_kpp_candidates_rows', _lib.ptr(X), _lib.ptr(x_sq), n, d, _lib.ptr(rows), 6, _lib.ptr(closest), _lib.ptr(dist), _lib.ptr(potl), _lib.ptr(ws), ws_bytes, st))
print(f'kpp_candidates:{t:.3f} ms')
