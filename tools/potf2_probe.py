import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
_lib.load()
for n in (32, 64, 96):
    batch = 148 * 64
    rng = np.random.default_rng(0)
    M = rng.normal(size=(n, n)); S = M @ M.T / n + np.eye(n)
    A0 = torch.as_tensor(S, device='cuda').repeat(batch, 1, 1).contiguous()
    info = torch.zeros(batch, dtype=torch.int32, device='cuda')
    best = 1e9
    for rep in range(3):
        A = A0.clone(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call('kmpx_potrf', _lib.ptr(A), n, n * n, None, n, batch, _lib.ptr(info), None, 0, _lib.stream())
        e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    per = best * 1e3 / (batch / 148 / (2 if n <= 300 else 1))
    print(f'n={n}: {best:.3f} ms for {batch} problems -> {per:.2f} us per problem per CTA slot')
