"""Per-phase cycle counters (CTA 0) of the small potrf / trtri kernels at n=404, TMA (flags 0) vs cp.async (flags 16)."""
import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
n, batch, lda = 404, int(os.environ.get('BATCH', 148)), 408
rng = np.random.default_rng(0)
M = rng.normal(size=(n, n)); S = M @ M.T / n + np.eye(n)
A0 = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A0[:, :, :n] = torch.as_tensor(S, device='cuda')
info = torch.zeros(batch, dtype=torch.int32, device='cuda'); buf = (ctypes.c_longlong * 16)()
extra = [int(a) for a in sys.argv[1:]]
for flags in [0, 16] + extra:
    lib.kmpx_debug_set(flags)
    for rep in range(2):
        A = A0.clone(); torch.cuda.synchronize(); lib.kmpx_debug_phases(buf, 1)
        _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, batch, _lib.ptr(info), None, 0, _lib.stream())
        _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, batch, None, 0, _lib.stream()); torch.cuda.synchronize()
    lib.kmpx_debug_phases(buf, 0)
    v = [x / 1e3 for x in buf]
    print(f'flags={flags}: potrf ' + ' '.join(f'p{i}={v[i]:.0f}k' for i in range(6)) + f' sum={sum(v[:6]):.0f}k | trtri ' +
          ' '.join(f'p{i}={v[i]:.0f}k' for i in range(8, 12)) + f' sum={sum(v[8:12]):.0f}k | aux wait={v[12]:.0f}k potf2={v[13]:.0f}k trtri={v[14]:.0f}k store={v[15]:.0f}k (potf2: 8x8 part={v[6]:.0f}k panel part={v[7]:.0f}k)')
lib.kmpx_debug_set(0)
