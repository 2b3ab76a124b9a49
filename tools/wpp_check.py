"""Correctness + timing of the warp-per-problem factor kernels (kmpx_potrf_wpp / kmpx_trtri_wpp) against numpy and
against the CTA-per-problem kernels.  usage: wpp_check.py [check] [time] [struct]"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(5)

def spd(n):
    M = rng.normal(size=(n, n)); return M @ M.T / n + np.eye(n)

def struct_spd(N, Ns):
    """component-major two-output system [[A00, D], [D, A11]] with diagonal D, padding rows identity"""
    n = 2 * Ns
    A = np.eye(n)
    t = np.sort(rng.uniform(0, 1, N))
    K = np.exp(-30.0 * (t[:, None] - t[None, :]) ** 2)
    for a in range(2):
        A[a * Ns:a * Ns + N, a * Ns:a * Ns + N] = K + np.diag(rng.uniform(0.5, 1.5, N))
    d = rng.uniform(-0.3, 0.3, N)
    A[Ns:Ns + N, :N] += np.diag(d); A[:N, Ns:Ns + N] += np.diag(d)
    return A

def check(sizes, nmax, lda, mats=None, struct_ns=0):
    A = torch.full((len(sizes), nmax, lda), float('nan'), dtype=torch.float64, device='cuda')
    mats = mats or [spd(n) for n in sizes]
    for b, n in enumerate(sizes):
        A[b, :n, :n] = torch.as_tensor(np.tril(mats[b]) + np.triu(np.full((n, n), np.nan), 1), device='cuda')
    n_sys = torch.as_tensor(np.array(sizes, dtype=np.int32), device='cuda'); info = torch.full((len(sizes),), -7, dtype=torch.int32, device='cuda')
    _lib.call('kmpx_potrf_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), _lib.ptr(info), struct_ns, _lib.stream())
    torch.cuda.synchronize()
    Lh = A.cpu().numpy(); worst = 0.0
    for b, n in enumerate(sizes):
        ref = np.linalg.cholesky(mats[b]); got = np.tril(Lh[b, :n, :n])
        e = np.abs(got - ref).max() / np.abs(ref).max(); worst = max(worst, e if np.isfinite(e) else 1e9)
    print(f'nmax={nmax} ns={struct_ns} potrf worst rel err {worst:.2e} info={sorted(set(info.cpu().numpy().tolist()))}')
    _lib.call('kmpx_trtri_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), struct_ns, _lib.stream())
    torch.cuda.synchronize()
    Xh = A.cpu().numpy(); worst = 0.0
    for b, n in enumerate(sizes):
        ref = np.linalg.inv(np.linalg.cholesky(mats[b])); got = np.tril(Xh[b, :n, :n])
        e = np.abs(got - ref).max() / np.abs(ref).max(); worst = max(worst, e if np.isfinite(e) else 1e9)
        if n < nmax: assert np.isnan(Xh[b, n:, :]).all(), 'rows beyond the system were touched'
    print(f'nmax={nmax} ns={struct_ns} trtri worst rel err {worst:.2e}')
    return Xh

def timing(n, batch, struct_ns=0, S=None, modes=('cta', 'wpp0', 'wpp128', 'wpp256')):
    lda = (n + 7) // 8 * 8
    S = spd(n) if S is None else S
    A0 = torch.zeros((batch, n, lda), dtype=torch.float64, device='cuda'); A0[:, :, :n] = torch.as_tensor(np.tril(S), device='cuda')
    info = torch.zeros(batch, dtype=torch.int32, device='cuda')
    A = torch.empty_like(A0)
    out = {}
    for mode in modes:
        lib.kmpx_debug_set(0 if mode == 'cta' else int(mode[3:]))
        rs = []
        for rep in range(4):
            A.copy_(A0); torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            if mode == 'cta':
                _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, batch, _lib.ptr(info), None, 0, _lib.stream())
                e[1].record()
                _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, batch, None, 0, _lib.stream())
            else:
                _lib.call('kmpx_potrf_wpp', _lib.ptr(A), lda, n * lda, None, n, batch, _lib.ptr(info), struct_ns, _lib.stream())
                e[1].record()
                _lib.call('kmpx_trtri_wpp', _lib.ptr(A), lda, n * lda, None, n, batch, struct_ns, _lib.stream())
            e[2].record(); e[2].synchronize()
            if rep: rs.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
        out[mode] = np.median(np.array(rs), axis=0)
        print(f'{mode} n={n} batch={batch} ns={struct_ns}: potrf {out[mode][0]:.2f} ms  trtri {out[mode][1]:.2f} ms', flush=True)
    lib.kmpx_debug_set(0)
    return out

if __name__ == '__main__':
    what = sys.argv[1:] or ['check', 'time']
    if 'time' in what: what = what[:what.index('time') + 1]
    if 'check' in what:
      for fl in (0, 128, 256):
        lib.kmpx_debug_set(fl); print('variant flags', fl)
        check([64, 65, 100, 257, 402, 404, 333, 96] * 5, 404, 408)
        check([64, 100, 257, 453, 512], 512, 512)
        check([200, 202] * 40, 202, 202)
        check([404] * 3000, 404, 408)
      lib.kmpx_debug_set(0)
    if 'struct' in what:
      for fl in (0, 128, 256):
        lib.kmpx_debug_set(fl); print('variant flags', fl)
        mats = [struct_spd(N, (N + 1) & ~1) for N in (202, 201, 200, 150)]
        sizes = [2 * ((N + 1) & ~1) for N in (202, 201, 200, 150)]
        for N, m in zip((202, 201, 200, 150), mats):
            Ns = (N + 1) & ~1
            x0 = check([2 * Ns], 2 * Ns, 2 * Ns + 4, [m], 0)
            x1 = check([2 * Ns], 2 * Ns, 2 * Ns + 4, [m], 2)
            print('struct vs plain bit-identical (lower):', np.array_equal(np.tril(x0[0, :, :2 * Ns]), np.tril(x1[0, :, :2 * Ns])))
    lib.kmpx_debug_set(0)
    if 'time' in what:
        S = struct_spd(202, 202)
        modes = sys.argv[sys.argv.index('time') + 1:] or ['cta', 'wpp0', 'wpp128', 'wpp256']
        timing(404, 5328, 2, S, modes=modes)
