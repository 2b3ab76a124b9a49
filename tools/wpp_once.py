"""One launch each of the warp-per-problem factor kernels on a cfg-3 sized chunk (for ncu).
usage: wpp_once.py [struct_ns] [batch] [debug flags ...]"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from wpp_check import struct_spd
lib = _lib.load()
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 5328
flags = [int(a) for a in sys.argv[3:]] or [0]
n, lda = 404, 408
S = struct_spd(202, 202)
A0 = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A0[:, :, :n] = torch.as_tensor(np.tril(S), device='cuda')
info = torch.zeros(batch, dtype=torch.int32, device='cuda')
for f in flags:
    lib.kmpx_debug_set(f)
    A = A0.clone()
    _lib.call('kmpx_potrf_wpp', _lib.ptr(A), lda, n * lda, None, n, batch, _lib.ptr(info), ns, _lib.stream())
    _lib.call('kmpx_trtri_wpp', _lib.ptr(A), lda, n * lda, None, n, batch, ns, _lib.stream())
    torch.cuda.synchronize()
lib.kmpx_debug_set(0)
print('ok', int(info.abs().sum()))
