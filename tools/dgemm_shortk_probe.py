"""k_dgemm on the shape of the blocked Cholesky's trailing update: C(m x m, lower tiles) -= P P^T with K = 256."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
m = 16128
C = torch.zeros((m, 16384), dtype=torch.float64, device='cuda')
for K, ldp, tri, beta in ((256, 256, 1, 1.0), (256, 256, 1, 0.0), (256, 256, 0, 1.0), (512, 512, 1, 1.0), (1024, 1024, 1, 1.0)):
    P = torch.randn((m, ldp), dtype=torch.float64, device='cuda')
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call('kmpx_dgemm', 0, 1, m, m, K, -1.0, _lib.ptr(P), ldp, 0, _lib.ptr(P), ldp, 0, beta, _lib.ptr(C), 16384, 0, 1, tri, _lib.stream())
        e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    fl = 2.0 * m * m * K * (0.5 if tri == 1 else 1.0)
    print(f'K={K} ldp={ldp} tri={tri} beta={beta}: {min(ts):.2f} ms  {fl / min(ts) * 1e-9:.2f} TFLOP/s')
