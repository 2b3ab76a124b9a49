"""One NT product at 4096^3 and one lower-tile rank-256 update at m = 8192 on the TMA-fed persistent GEMM (for ncu captures
of k_dgemm_nt_tma)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
n = 4096
A = torch.randn((n, n), dtype=torch.float64, device='cuda'); B = torch.randn((n, n), dtype=torch.float64, device='cuda'); C = torch.zeros((n, n), dtype=torch.float64, device='cuda')
m = 8192
P = torch.randn((m, 256), dtype=torch.float64, device='cuda'); S = torch.zeros((m, m), dtype=torch.float64, device='cuda')
for rep in range(2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    _lib.call('kmpx_dgemm', 0, 1, n, n, n, 1.0, _lib.ptr(A), n, 0, _lib.ptr(B), n, 0, 0.0, _lib.ptr(C), n, 0, 1, 0, _lib.stream())
    e[1].record()
    _lib.call('kmpx_dgemm', 0, 1, m, m, 256, -1.0, _lib.ptr(P), 256, 0, _lib.ptr(P), 256, 0, 1.0, _lib.ptr(S), m, 0, 1, 1, _lib.stream())
    e[2].record(); e[2].synchronize()
    print(f'NT {n}^3: {e[0].elapsed_time(e[1]):.3f} ms ({2.0 * n ** 3 / e[0].elapsed_time(e[1]) * 1e-9:.2f} TF); lower rank-256 update m={m}: '
          f'{e[1].elapsed_time(e[2]):.3f} ms ({m * m * 256.0 / e[1].elapsed_time(e[2]) * 1e-9:.2f} TF)')
