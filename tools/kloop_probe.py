import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
n, batch = 404, 148
rng = np.random.default_rng(0)
M = rng.normal(size=(n, n)); S = M @ M.T / n + np.eye(n)
lda = 408
A0 = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A0[:, :, :n] = torch.as_tensor(S, device='cuda')
info = torch.zeros(batch, dtype=torch.int32, device='cuda'); buf = (ctypes.c_longlong * 16)()
for flags, name in ((0, 'normal'), (1, 'no MMA'), (2, 'no loads'), (3, 'neither')):
    lib.kmpx_debug_set(flags)
    for rep in range(2):
        A = A0.clone(); torch.cuda.synchronize(); lib.kmpx_debug_phases(buf, 1)
        _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, batch, _lib.ptr(info), None, 0, _lib.stream()); torch.cuda.synchronize()
    lib.kmpx_debug_phases(buf, 0)
    print(f'{name:10s} potrf k_loop cycles = {buf[0]/1e3:.1f}k')
lib.kmpx_debug_set(0)
