"""A/B timing of kmpx_potrf / kmpx_trtri on a cfg-3 sized chunk (4440 systems, n = 404) with debug flag sets given
on the command line (default: 0 vs 32 = no L2 prefetch), alternating, CUDA events."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
n, batch, lda = 404, 4440, 408
flagsets = [int(a) for a in sys.argv[1:]] or [0, 32]
rng = np.random.default_rng(0)
M = rng.normal(size=(n, n)); S = M @ M.T / n + np.eye(n)
A0 = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A0[:, :, :n] = torch.as_tensor(S, device='cuda')
A = torch.empty_like(A0)
info = torch.zeros(batch, dtype=torch.int32, device='cuda')
res = {f: [[], []] for f in flagsets}
for rep in range(6):
    for f in flagsets:
        lib.kmpx_debug_set(f)
        A.copy_(A0); torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, batch, _lib.ptr(info), None, 0, _lib.stream())
        e[1].record()
        _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, batch, None, 0, _lib.stream())
        e[2].record(); torch.cuda.synchronize()
        if rep:
            res[f][0].append(e[0].elapsed_time(e[1])); res[f][1].append(e[1].elapsed_time(e[2]))
lib.kmpx_debug_set(0)
for f in flagsets:
    print(f'flags={f}: potrf {np.median(res[f][0]):.3f} ms  trtri {np.median(res[f][1]):.3f} ms  (min {min(res[f][0]):.3f} / {min(res[f][1]):.3f})')
