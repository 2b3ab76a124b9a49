"""Correctness + timing of kmpx_potrf / kmpx_trtri on ragged small batches (TMA and cp.async variants)."""
import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(5)

def spd(n):
    M = rng.normal(size=(n, n)); return M @ M.T / n + np.eye(n)

def check(sizes, nmax, lda, flags):
    lib.kmpx_debug_set(flags)
    A = torch.full((len(sizes), nmax, lda), float('nan'), dtype=torch.float64, device='cuda')
    mats = []
    for b, n in enumerate(sizes):
        S = spd(n); mats.append(S)
        A[b, :n, :n] = torch.as_tensor(np.tril(S) + np.triu(np.full((n, n), np.nan), 1), device='cuda')
    n_sys = torch.as_tensor(np.array(sizes, dtype=np.int32), device='cuda'); info = torch.zeros(len(sizes), dtype=torch.int32, device='cuda')
    _lib.call('kmpx_potrf', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), _lib.ptr(info), None, 0, _lib.stream())
    torch.cuda.synchronize()
    Lh = A.cpu().numpy(); worst = 0.0
    for b, n in enumerate(sizes):
        ref = np.linalg.cholesky(mats[b]); got = np.tril(Lh[b, :n, :n])
        e = np.abs(got - ref).max() / np.abs(ref).max(); worst = max(worst, e if np.isfinite(e) else 1e9)
    print(f'flags={flags} nmax={nmax} potrf worst rel err {worst:.2e} info={info.cpu().numpy().tolist()}')
    _lib.call('kmpx_trtri', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), None, 0, _lib.stream())
    torch.cuda.synchronize()
    Xh = A.cpu().numpy(); worst = 0.0
    for b, n in enumerate(sizes):
        ref = np.linalg.inv(np.linalg.cholesky(mats[b])); got = np.tril(Xh[b, :n, :n])
        e = np.abs(got - ref).max() / np.abs(ref).max(); worst = max(worst, e if np.isfinite(e) else 1e9)
    print(f'flags={flags} nmax={nmax} trtri worst rel err {worst:.2e}')
    lib.kmpx_debug_set(0)

def timing(n, batch, flags):
    lib.kmpx_debug_set(flags)
    lda = (n + 7) // 8 * 8
    S = spd(n)
    A0 = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A0[:, :, :n] = torch.as_tensor(S, device='cuda')
    info = torch.zeros(batch, dtype=torch.int32, device='cuda')
    res = []
    for rep in range(3):
        A = A0.clone(); torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, batch, _lib.ptr(info), None, 0, _lib.stream())
        e[1].record()
        _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, batch, None, 0, _lib.stream())
        e[2].record(); e[2].synchronize()
        res = (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]))
    print(f'flags={flags} n={n} batch={batch}: potrf {res[0]:.2f} ms  trtri {res[1]:.2f} ms')
    lib.kmpx_debug_set(0)

if __name__ == '__main__':
    for flags in (0, 16):
        check([64, 65, 100, 257, 402, 404, 333, 96], 404, 408, flags)
        check([64, 100, 257, 453, 512], 512, 512, flags)
        check([200, 202], 202, 202, flags)
    for flags in (0, 16):
        timing(404, 2048, flags)
        timing(202, 2048, flags)
