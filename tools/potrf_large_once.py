"""One kmpx_potrf + kmpx_trtri at n = 16384 (for ncu launch lists / timing)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device='cuda'); g.manual_seed(0)
M = torch.randn((n, n), dtype=torch.float64, device='cuda', generator=g)
A0 = (M @ M.T) / n + torch.eye(n, dtype=torch.float64, device='cuda') * 2.0
del M
nbytes = max(int(lib.kmpx_potrf_workspace_bytes(n, 1)), int(lib.kmpx_trtri_workspace_bytes(n, 1)))
ws = torch.empty(nbytes // 8, dtype=torch.float64, device='cuda')
info = torch.zeros(1, dtype=torch.int32, device='cuda')
lib.kmpx_debug_set(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
for rep in range(3):
    A = A0.clone(); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    _lib.call('kmpx_potrf', _lib.ptr(A), n, n * n, None, n, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, _lib.stream())
    e[1].record()
    _lib.call('kmpx_trtri', _lib.ptr(A), n, n * n, None, n, 1, _lib.ptr(ws), nbytes, _lib.stream())
    e[2].record(); e[2].synchronize()
    print(f'n={n}: potrf {e[0].elapsed_time(e[1]):.2f} ms ({n**3/3/e[0].elapsed_time(e[1])*1e-9:.2f} TF)  trtri {e[1].elapsed_time(e[2]):.2f} ms  info={int(info.item())}')
