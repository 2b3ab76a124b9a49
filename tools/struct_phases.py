"""Per-phase cycle counters (CTA 0) of the fused fit kernel k_factor_struct at the cfg-3 shape (N = 202, O = 2)."""
import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
N, O, B = 202, 2, int(os.environ.get('BATCH', 148 * 8))
n = 2 * N; lda = 408
rng = np.random.default_rng(0)
s = np.tile((0.01 * np.arange(1, N + 1)).reshape(1, N, 1), (B, 1, 1))
Bm = rng.normal(size=(B, N, O, O)) * 0.3
sig = np.einsum('bnac,bndc->bnad', Bm, Bm) + 0.05 * np.eye(O)
sd, gd = torch.as_tensor(s, device='cuda'), torch.as_tensor(sig, device='cuda')
F = torch.empty((B, n, lda), dtype=torch.float64, device='cuda'); info = torch.zeros(B, dtype=torch.int32, device='cuda')
buf = (ctypes.c_longlong * 16)()
for rep in range(3):
    torch.cuda.synchronize(); lib.kmpx_debug_phases(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call('kmpx_factor_struct', _lib.ptr(sd), _lib.ptr(gd), None, _lib.ptr(F), lda, n * lda, N, 1, O, B, 200.0, 0.5, _lib.ptr(info), _lib.stream())
    e1.record(); e1.synchronize()
lib.kmpx_debug_phases(buf, 0)
per = B // 148
v = [x / 1e3 / per for x in buf]
names = {0: 'build', 1: 'chol diag', 2: 'chol panel', 3: 'chol trailing', 4: 'trtri diag', 5: 'trtri Tm', 8: 'trtri out', 9: 'write / misc', 10: 'G,W,S', 11: 'X21'}
print(f'{e0.elapsed_time(e1):.3f} ms for {B} problems = {e0.elapsed_time(e1) * 1e3 / per:.1f} us per problem per SM; kcycles per problem: ' +
      ', '.join(f'{names[i]} {v[i]:.1f}' for i in names) + f'; sum {sum(v[i] for i in names):.0f} (aux 8x8 {v[6]:.0f}, aux panel {v[7]:.0f}); info max {int(info.max())}')
