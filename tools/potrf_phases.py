"""Cycle breakdown of k_potrf_small per phase (CTA 0): K-loop, panel load, potf2, trtri32+copy, chunk mul, store."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
names = ['k_loop', 'panel_load', 'potf2', 'trtri32+copy', 'chunk_mul', 'store']
for n, batch in ((32, 148), (64, 148), (404, 148)):
    rng = np.random.default_rng(0)
    M = rng.normal(size=(n, n)); S = M @ M.T / n + np.eye(n)
    lda = (n + 7) // 8 * 8
    A = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A[:, :, :n] = torch.as_tensor(S, device='cuda')
    info = torch.zeros(batch, dtype=torch.int32, device='cuda')
    buf = (ctypes.c_longlong * 16)()
    A0 = A.clone()
    for rep in range(2):
        A.copy_(A0); torch.cuda.synchronize()
        lib.kmpx_debug_phases(buf, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, batch, _lib.ptr(info), None, 0, _lib.stream())
        e1.record(); e1.synchronize()
    lib.kmpx_debug_phases(buf, 0)
    v = np.array(list(buf)[:6], dtype=float)
    print(f'n={n} batch={batch}: {e0.elapsed_time(e1)*1e3:.0f} us total; cycles/CTA0: ' + ', '.join(f'{a}={b/1e3:.1f}k ({b/v.sum():.0%})' for a, b in zip(names, v)))

tnames = ['diag load+trtri32', 'T=L21*Dinv', 'k_loop', 'store']
for n, batch in ((404, 148), (256, 1)):
    rng = np.random.default_rng(0)
    M = rng.normal(size=(n, n)); S = M @ M.T / n + np.eye(n)
    lda = (n + 7) // 8 * 8
    Lh = np.linalg.cholesky(S)
    A = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A[:, :, :n] = torch.as_tensor(Lh, device='cuda')
    A0 = A.clone(); buf = (ctypes.c_longlong * 16)()
    for rep in range(2):
        A.copy_(A0); torch.cuda.synchronize(); lib.kmpx_debug_phases(buf, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, batch, None, 0, _lib.stream()); e1.record(); e1.synchronize()
    lib.kmpx_debug_phases(buf, 0)
    v = np.array(list(buf)[8:12], dtype=float)
    print(f'trtri n={n} batch={batch}: {e0.elapsed_time(e1)*1e3:.0f} us total; cycles/CTA0: ' + ', '.join(f'{a}={b/1e3:.1f}k ({b/v.sum():.0%})' for a, b in zip(tnames, v)))
