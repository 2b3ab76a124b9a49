"""Is the streamed K-loop DRAM-bound or TMA-bound?  All CTAs factor the SAME matrix (strideA = 0 -> L2 resident)."""
import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
n, lda = 404, 408
rng = np.random.default_rng(0)
M = rng.normal(size=(n, n)); S = M @ M.T / n + np.eye(n)
buf = (ctypes.c_longlong * 16)()
for batch, stride in ((148, n * lda), (148, 0), (16, n * lda)):
    A0 = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A0[:, :, :n] = torch.as_tensor(S, device='cuda')
    info = torch.zeros(batch, dtype=torch.int32, device='cuda')
    for flags in (1, 17):
        lib.kmpx_debug_set(flags)
        for rep in range(2):
            A = A0.clone(); torch.cuda.synchronize(); lib.kmpx_debug_phases(buf, 1)
            _lib.call('kmpx_potrf', _lib.ptr(A), lda, stride, None, n, batch, _lib.ptr(info), None, 0, _lib.stream()); torch.cuda.synchronize()
        lib.kmpx_debug_phases(buf, 0)
        print(f'batch={batch} stride={stride} flags={flags}: k_loop(no MMA)={buf[0]/1e3:.0f}k p1={buf[1]/1e3:.0f}k')
lib.kmpx_debug_set(0)
