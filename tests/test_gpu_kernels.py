"""GPU parity of the individual kernels, called through the C ABI (kmp_demos_b200._lib), against numpy and the
CPU oracle.  Tolerances: bit-exact for labels/indices; fp64 kernels within the stated relative error."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def T():
    import torch
    from kmp_demos_b200 import _lib
    _lib.load()
    return torch


def dev(torch, a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), device='cuda') if dtype is None else \
        torch.as_tensor(np.ascontiguousarray(a), device='cuda', dtype=dtype)


def comp_perm(N, O):
    """index map point-major (i*O+a) -> component-major (a*Ns+i) and the padded size"""
    Ns = (N + 1) // 2 * 2
    idx = np.array([[a * Ns + i for a in range(O)] for i in range(N)]).reshape(-1)
    return idx, O * Ns


def test_kernel_build_matches_reference_matrix(T, golden):
    from kmp_demos_b200 import model, _lib
    z = golden('small_cases.npz')
    for c in range(int(z['kmp_cases'])):
        I, O, N, td, sf, lam, alpha = z[f'kmp{c}_cfg']
        O, N, td = int(O), int(N), bool(td)
        s, sig, Aref = z[f'kmp{c}_s'], z[f'kmp{c}_sigma'], z[f'kmp{c}_A']
        A = model.build_matrix(s, sig, lam, sf, td, _lib.LAYOUT_POINT)
        tol = 1e-10 if td else 1e-15      # FD blocks amplify 1-ulp exp differences by 1/dt^2
        assert rel_err(A, Aref) < tol, f'case {c} point-major'
        Ac = model.build_matrix(s, sig, lam, sf, td, _lib.LAYOUT_COMPONENT)
        idx, n = comp_perm(N, O)
        assert Ac.shape == (n, n)
        assert rel_err(Ac[np.ix_(idx, idx)], Aref) < tol, f'case {c} component-major'
        pad = np.setdiff1d(np.arange(n), idx)
        if pad.size:
            assert np.array_equal(Ac[np.ix_(pad, pad)], np.eye(pad.size)) and np.all(Ac[np.ix_(pad, idx)] == 0)


def test_kernel_build_lower_only_and_batch(T):
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(0)
    B, Nmax, O, I = 5, 37, 2, 1
    npts = np.array([37, 36, 1, 20, 33], dtype=np.int32)
    s = np.sort(rng.random((B, Nmax, I)), axis=1)
    Bm = rng.normal(size=(B, Nmax, O, O))
    sig = np.einsum('bnij,bnkj->bnik', Bm, Bm) + 0.1 * np.eye(O)
    n = O * ((Nmax + 1) // 2 * 2)
    lda = (n + 7) // 8 * 8
    A_full = T.full((B, n, lda), -7.0, dtype=T.float64, device='cuda')
    A_low = T.full((B, n, lda), -7.0, dtype=T.float64, device='cuda')
    s_d, sig_d, n_d = dev(T, s), dev(T, sig), dev(T, npts)       # keep the device tensors alive across the launches
    for A, uplo in ((A_full, _lib.UPLO_FULL), (A_low, _lib.UPLO_LOWER)):
        _lib.call('kmpx_kernel_build', _lib.ptr(s_d), _lib.ptr(sig_d), _lib.ptr(n_d), _lib.ptr(A), B, Nmax,
                  I, O, lda, n * lda, 200.0, 0.5, 0, _lib.LAYOUT_COMPONENT, uplo, _lib.stream())
    Af, Al = A_full.cpu().numpy(), A_low.cpu().numpy()
    from oracle import kmp_oracle as orc
    for b in range(B):
        N = int(npts[b])
        idx, nb = comp_perm(N, O)
        ref = orc.kmp_matrix(s[b, :N].T, np.transpose(sig[b, :N], (1, 2, 0)), 0.5, 200.0, False)
        assert rel_err(Af[b][np.ix_(idx, idx)], ref) < 1e-15
        low = np.tril(Af[b][:nb, :nb])
        assert np.array_equal(np.tril(Al[b][:nb, :nb]), low)
        assert np.all(Al[b][:nb, :nb][np.triu_indices(nb, 1)] == -7.0), 'LOWER must not touch the strict upper triangle'
        assert np.all(Af[b][nb:, :] == -7.0), 'rows beyond the system must stay untouched'


@pytest.mark.parametrize('ta,tb', [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_dgemm(T, ta, tb):
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(1)
    for (m, n, k) in [(128, 128, 64), (130, 70, 33), (257, 300, 129), (1, 5, 3)]:
        A = rng.normal(size=(k, m) if ta else (m, k))
        Bm = rng.normal(size=(n, k) if tb else (k, n))
        C0 = rng.normal(size=(m, n))
        lda = (A.shape[1] + 1) // 2 * 2
        ldb = (Bm.shape[1] + 1) // 2 * 2
        ldc = (n + 1) // 2 * 2
        Ad = T.zeros((A.shape[0], lda), dtype=T.float64, device='cuda'); Ad[:, :A.shape[1]] = dev(T, A)
        Bd = T.zeros((Bm.shape[0], ldb), dtype=T.float64, device='cuda'); Bd[:, :Bm.shape[1]] = dev(T, Bm)
        Cd = T.zeros((m, ldc), dtype=T.float64, device='cuda'); Cd[:, :n] = dev(T, C0)
        _lib.call('kmpx_dgemm', ta, tb, m, n, k, 1.5, _lib.ptr(Ad), lda, 0, _lib.ptr(Bd), ldb, 0, -0.5, _lib.ptr(Cd), ldc, 0, 1, 0,
                  _lib.stream())
        ref = 1.5 * (A.T if ta else A) @ (Bm.T if tb else Bm) - 0.5 * C0
        assert rel_err(Cd[:, :n].cpu().numpy(), ref) < 1e-13, (m, n, k)


def test_dgemm_triangular_modes(T):
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(2)
    m = 300
    A = rng.normal(size=(m, 64)); C0 = rng.normal(size=(m, m))
    Ad, Cd = dev(T, A), dev(T, C0)
    _lib.call('kmpx_dgemm', 0, 1, m, m, 64, -1.0, _lib.ptr(Ad), 64, 0, _lib.ptr(Ad), 64, 0, 1.0, _lib.ptr(Cd), m, 0, 1, 1, _lib.stream())
    out = Cd.cpu().numpy()
    assert rel_err(np.tril(out), np.tril(C0 - A @ A.T)) < 1e-13
    Lw = np.tril(rng.normal(size=(m, m))); Bm = rng.normal(size=(m, 40))
    Cd = T.zeros((m, 40), dtype=T.float64, device='cuda')
    Lw_d, Bm_d = dev(T, Lw), dev(T, Bm)
    _lib.call('kmpx_dgemm', 0, 0, m, 40, m, 1.0, _lib.ptr(Lw_d), m, 0, _lib.ptr(Bm_d), 40, 0, 0.0, _lib.ptr(Cd), 40, 0, 1, 2,
              _lib.stream())
    assert rel_err(Cd.cpu().numpy(), Lw @ Bm) < 1e-13
    Bl = np.tril(rng.normal(size=(m, m)))
    Cd = T.zeros((40, m), dtype=T.float64, device='cuda')
    A2 = rng.normal(size=(40, m))
    A2_d, Bl_d = dev(T, A2), dev(T, Bl)
    _lib.call('kmpx_dgemm', 0, 0, 40, m, m, 1.0, _lib.ptr(A2_d), m, 0, _lib.ptr(Bl_d), m, 0, 0.0, _lib.ptr(Cd), m, 0, 1, 3,
              _lib.stream())
    assert rel_err(Cd.cpu().numpy(), A2 @ Bl) < 1e-13


def test_dgemm_nt_tma_persistent_batched_and_triangular(T):
    """The TMA-fed persistent NT kernel (dgemm_tma.cu): more tiles than SMs, ragged edges, ring wrap-around (k = 300),
    strided batch, lower-tile mode (tri = 1) and triangular-operand mode (tri = 2), beta = 0 and beta != 0; the debug
    flag 128 routes the same call through the cp.async kernel, which must agree to rounding."""
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(7)
    g = T.Generator(device='cuda'); g.manual_seed(5)
    m, n, k = 2065, 1930, 300
    A = T.randn((m, k), generator=g, dtype=T.float64, device='cuda'); B = T.randn((n, k), generator=g, dtype=T.float64, device='cuda')
    C0 = T.randn((m, n), generator=g, dtype=T.float64, device='cuda')
    C = C0.clone()
    _lib.call('kmpx_dgemm', 0, 1, m, n, k, 0.75, _lib.ptr(A), k, 0, _lib.ptr(B), k, 0, 2.0, _lib.ptr(C), n, 0, 1, 0, _lib.stream())
    ref = 0.75 * A @ B.T + 2.0 * C0
    assert float((C - ref).abs().max() / ref.abs().max()) < 1e-13
    lib.kmpx_debug_set(128)
    C2 = C0.clone()
    _lib.call('kmpx_dgemm', 0, 1, m, n, k, 0.75, _lib.ptr(A), k, 0, _lib.ptr(B), k, 0, 2.0, _lib.ptr(C2), n, 0, 1, 0, _lib.stream())
    lib.kmpx_debug_set(0)
    assert float((C - C2).abs().max() / ref.abs().max()) < 1e-14
    # strided batch, beta = 0
    nb, mb, kb = 5, 260, 72
    Ab = T.randn((nb, mb, kb), generator=g, dtype=T.float64, device='cuda'); Bb = T.randn((nb, 200, kb), generator=g, dtype=T.float64, device='cuda')
    Cb = T.full((nb, mb, 200), float('nan'), dtype=T.float64, device='cuda')
    _lib.call('kmpx_dgemm', 0, 1, mb, 200, kb, 1.0, _lib.ptr(Ab), kb, mb * kb, _lib.ptr(Bb), kb, 200 * kb, 0.0, _lib.ptr(Cb), 200, mb * 200, nb, 0,
              _lib.stream())
    refb = T.bmm(Ab, Bb.transpose(1, 2))
    assert float((Cb - refb).abs().max() / refb.abs().max()) < 1e-13
    # tri = 1: lower tiles of a symmetric rank-k update, rectangular tail (m > n)
    m1, n1 = 1500, 1100
    P = T.randn((m1, 256), generator=g, dtype=T.float64, device='cuda')
    S0 = T.randn((m1, n1), generator=g, dtype=T.float64, device='cuda'); S = S0.clone()
    _lib.call('kmpx_dgemm', 0, 1, m1, n1, 256, -1.0, _lib.ptr(P), 256, 0, _lib.ptr(P), 256, 0, 1.0, _lib.ptr(S), n1, 0, 1, 1, _lib.stream())
    refs = S0 - P @ P[:n1].T
    low = T.tril(T.ones((m1, n1), dtype=T.bool, device='cuda'))
    assert float(((S - refs) * low).abs().max() / refs.abs().max()) < 1e-13
    # tri = 2: lower-triangular A (explicit zeros above the diagonal), k range cut at the diagonal tile
    mt = 1000
    Lw = T.tril(T.randn((mt, mt), generator=g, dtype=T.float64, device='cuda'))
    Kt = T.randn((384, mt), generator=g, dtype=T.float64, device='cuda')
    V = T.empty((mt, 384), dtype=T.float64, device='cuda')
    _lib.call('kmpx_dgemm', 0, 1, mt, 384, mt, 1.0, _lib.ptr(Lw), mt, 0, _lib.ptr(Kt), mt, 0, 0.0, _lib.ptr(V), 384, 0, 1, 2, _lib.stream())
    refv = Lw @ Kt.T
    assert float((V - refv).abs().max() / refv.abs().max()) < 1e-13


def _spd(rng, n):
    M = rng.normal(size=(n, n))
    return M @ M.T / n + np.eye(n)


@pytest.mark.parametrize('O', [1, 2])
def test_factor_struct_equals_potrf_trtri(T, O):
    """The fused shared-memory-resident fit (k_factor_struct: build + Cholesky + inverse exploiting the diagonal
    cross-component blocks) leaves the same L^-1 (component-major, lower triangle) as kmpx_kernel_build + kmpx_potrf +
    kmpx_trtri, on a ragged batch including odd N (identity padding rows) and tiny systems; a non-SPD problem is
    reported through info."""
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11 + O)
    sizes = [202, 201, 200, 33, 64, 1, 7, 100, 204, 31, 32, 65, 129, 193]
    B, nmax = len(sizes), 204
    assert lib.kmpx_factor_struct_applicable(nmax, 1, O) == 1
    n_sys_max = int(lib.kmpx_sys_size(nmax, O, _lib.LAYOUT_COMPONENT))
    lda = (n_sys_max + 7) // 8 * 8
    s = np.zeros((B, nmax, 1)); sig = np.zeros((B, nmax, O, O))
    for b, N in enumerate(sizes):
        s[b, :N, 0] = 0.01 * np.arange(1, N + 1) + 0.002 * rng.random(N)
        Bm = rng.normal(size=(N, O, O)) * 0.3
        sig[b, :N] = np.einsum('nab,ncb->nac', Bm, Bm) + 0.05 * np.eye(O)
    sd, gd = dev(T, s), dev(T, sig)
    npts = dev(T, np.array(sizes, dtype=np.int32))
    st = _lib.stream()
    A = T.full((B, n_sys_max, lda), float('nan'), dtype=T.float64, device='cuda')
    _lib.call('kmpx_kernel_build', _lib.ptr(sd), _lib.ptr(gd), _lib.ptr(npts), _lib.ptr(A), B, nmax, 1, O, lda, n_sys_max * lda,
              200.0, 0.5, 0, _lib.LAYOUT_COMPONENT, _lib.UPLO_LOWER, st)
    nsys = T.empty(B, dtype=T.int32, device='cuda')
    _lib.call('kmpx_sys_sizes', _lib.ptr(npts), O, _lib.LAYOUT_COMPONENT, B, _lib.ptr(nsys), st)
    info = T.full((B,), -1, dtype=T.int32, device='cuda')
    _lib.call('kmpx_potrf', _lib.ptr(A), lda, n_sys_max * lda, _lib.ptr(nsys), n_sys_max, B, _lib.ptr(info), None, 0, st)
    _lib.call('kmpx_trtri', _lib.ptr(A), lda, n_sys_max * lda, _lib.ptr(nsys), n_sys_max, B, None, 0, st)
    assert int(info.abs().max().item()) == 0
    F = T.full((B, n_sys_max, lda), float('nan'), dtype=T.float64, device='cuda')
    info2 = T.full((B,), -1, dtype=T.int32, device='cuda')
    _lib.call('kmpx_factor_struct', _lib.ptr(sd), _lib.ptr(gd), _lib.ptr(npts), _lib.ptr(F), lda, n_sys_max * lda, nmax, 1, O, B,
              200.0, 0.5, _lib.ptr(info2), st)
    assert int(info2.abs().max().item()) == 0
    Ah, Fh = A.cpu().numpy(), F.cpu().numpy()
    for b, N in enumerate(sizes):
        n = int(nsys[b].item())
        ref, got = np.tril(Ah[b, :n, :n]), np.tril(Fh[b, :n, :n])
        assert np.all(np.isfinite(got)), (N, np.argwhere(~np.isfinite(got))[:4])
        assert np.max(np.abs(got - ref)) < 1e-11 * np.max(np.abs(ref)), (N, float(np.max(np.abs(got - ref))), float(np.max(np.abs(ref))))
    # a system that is not positive definite (negative covariance block) is flagged, not silently factorised
    sig_bad = sig.copy(); sig_bad[0, 5] = -50.0 * np.eye(O)
    _lib.call('kmpx_factor_struct', _lib.ptr(sd), _lib.ptr(dev(T, sig_bad)), _lib.ptr(npts), _lib.ptr(F), lda, n_sys_max * lda, nmax, 1, O, B,
              200.0, 0.5, _lib.ptr(info2), st)
    ih = info2.cpu().numpy()
    assert ih[0] > 0 and np.all(ih[1:] == 0)


def test_potrf_trtri_small_ragged_batch(T):
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(3)
    sizes = [1, 5, 31, 32, 33, 64, 100, 257, 402, 404, 453, 512]
    nmax = 512
    lda = nmax
    A = T.full((len(sizes), nmax, lda), float('nan'), dtype=T.float64, device='cuda')
    mats = []
    for b, n in enumerate(sizes):
        S = _spd(rng, n)
        mats.append(S)
        low = np.tril(S) + np.triu(np.full((n, n), np.nan), 1)     # strict upper = NaN: must never be read
        A[b, :n, :n] = dev(T, low)
    n_sys = dev(T, np.array(sizes, dtype=np.int32))
    info = T.full((len(sizes),), -1, dtype=T.int32, device='cuda')
    _lib.call('kmpx_potrf', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), _lib.ptr(info), None, 0, _lib.stream())
    L = A.cpu().numpy()
    assert np.all(info.cpu().numpy() == 0)
    for b, n in enumerate(sizes):
        assert rel_err(np.tril(L[b, :n, :n]), np.linalg.cholesky(mats[b])) < 1e-13, n
    _lib.call('kmpx_trtri', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), None, 0, _lib.stream())
    Li = A.cpu().numpy()
    for b, n in enumerate(sizes):
        ref = np.linalg.inv(np.linalg.cholesky(mats[b]))
        assert rel_err(np.tril(Li[b, :n, :n]), ref) < 1e-12, n


@pytest.mark.parametrize('nmax,lda,sizes', [(404, 408, [64, 65, 96, 100, 257, 333, 402, 404, 1, 5, 31, 32, 33]), (202, 202, [200, 202, 128]),
                                            (480, 480, [480, 449, 33]),
                                            (512, 512, [512, 511, 497, 481, 480, 300])])      # n > 480: the fifteenth 32-row unit (warp 7)
def test_potrf_trtri_tma_variants_ragged(T, nmax, lda, sizes):
    """Sizes at which BOTH small kernels take the TMA path (64 <= n_max, buffers fit): ragged batches, NaN everywhere
    outside the lower triangles, leading dimension with and without padding columns."""
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(11)
    A = T.full((len(sizes), nmax, lda), float('nan'), dtype=T.float64, device='cuda')
    mats = []
    for b, n in enumerate(sizes):
        S = _spd(rng, n)
        mats.append(S)
        A[b, :n, :n] = dev(T, np.tril(S) + np.triu(np.full((n, n), np.nan), 1))
    n_sys = dev(T, np.array(sizes, dtype=np.int32))
    info = T.zeros(len(sizes), dtype=T.int32, device='cuda')
    _lib.call('kmpx_potrf', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), _lib.ptr(info), None, 0, _lib.stream())
    Lh = A.cpu().numpy()
    assert info.cpu().numpy().tolist() == [0] * len(sizes)
    for b, n in enumerate(sizes):
        assert rel_err(np.tril(Lh[b, :n, :n]), np.linalg.cholesky(mats[b])) < 1e-13, n
    _lib.call('kmpx_trtri', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), None, 0, _lib.stream())
    Xh = A.cpu().numpy()
    for b, n in enumerate(sizes):
        assert rel_err(np.tril(Xh[b, :n, :n]), np.linalg.inv(np.linalg.cholesky(mats[b]))) < 1e-12, n
        assert np.all(np.isnan(Xh[b, n:, :])), 'rows beyond the system must stay untouched'


def test_potrf_tma_reports_first_bad_pivot(T):
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(12)
    n = 160
    good = _spd(rng, n)
    bad = good.copy(); bad[100, 100] = -5.0
    A = dev(T, np.stack([good, bad, good]))
    info = T.zeros(3, dtype=T.int32, device='cuda')
    _lib.call('kmpx_potrf', _lib.ptr(A), n, n * n, None, n, 3, _lib.ptr(info), None, 0, _lib.stream())
    assert info.cpu().numpy().tolist() == [0, 101, 0]
    assert rel_err(np.tril(A[2].cpu().numpy()), np.linalg.cholesky(good)) < 1e-13


def _struct_spd(rng, N):
    """Component-major two-output KMP system [[A00, D], [D, A11]] with diagonal D (padding rows identity): what
    kmpx_kernel_build writes for O = 2 and the plain kernel."""
    Ns = (N + 1) // 2 * 2
    A = np.eye(2 * Ns)
    t = np.sort(rng.uniform(0, 1, N))
    K = np.exp(-30.0 * (t[:, None] - t[None, :]) ** 2)
    for a in range(2):
        A[a * Ns:a * Ns + N, a * Ns:a * Ns + N] = K + np.diag(rng.uniform(0.5, 1.5, N))
    d = rng.uniform(-0.3, 0.3, N)
    A[Ns:Ns + N, :N] += np.diag(d)
    A[:N, Ns:Ns + N] += np.diag(d)
    return A


@pytest.mark.parametrize('variant_flags', [0, 128, 256])
@pytest.mark.parametrize('nmax,lda,sizes', [(404, 408, [64, 65, 96, 100, 257, 333, 402, 404, 1, 5, 31, 32, 33] * 3),
                                            (202, 202, [200, 202, 128]), (512, 512, [512, 511, 497, 481, 480, 300, 64])])
def test_potrf_trtri_warp_per_problem_ragged(T, nmax, lda, sizes, variant_flags):
    """kmpx_potrf_wpp / kmpx_trtri_wpp (factor_wpp.cuh; all three tile variants) on ragged batches with NaN everywhere
    outside the lower triangles: factor and inverse against numpy, rows beyond a system untouched."""
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(21)
    A = T.full((len(sizes), nmax, lda), float('nan'), dtype=T.float64, device='cuda')
    mats = []
    for b, n in enumerate(sizes):
        S = _spd(rng, n)
        mats.append(S)
        A[b, :n, :n] = dev(T, np.tril(S) + np.triu(np.full((n, n), np.nan), 1))
    n_sys = dev(T, np.array(sizes, dtype=np.int32))
    info = T.full((len(sizes),), -3, dtype=T.int32, device='cuda')
    assert lib.kmpx_factor_wpp_applicable(_lib.ptr(A), lda, nmax * lda, nmax, len(sizes)) == 1
    lib.kmpx_debug_set(variant_flags)
    try:
        _lib.call('kmpx_potrf_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), _lib.ptr(info), 0, _lib.stream())
        Lh = A.cpu().numpy()
        _lib.call('kmpx_trtri_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), 0, _lib.stream())
        Xh = A.cpu().numpy()
    finally:
        lib.kmpx_debug_set(0)
    assert info.cpu().numpy().tolist() == [0] * len(sizes)
    for b, n in enumerate(sizes):
        assert rel_err(np.tril(Lh[b, :n, :n]), np.linalg.cholesky(mats[b])) < 1e-13, n
        assert rel_err(np.tril(Xh[b, :n, :n]), np.linalg.inv(np.linalg.cholesky(mats[b]))) < 1e-12, n
        assert np.all(np.isnan(Xh[b, n:, :])), 'rows beyond the system must stay untouched'


@pytest.mark.parametrize('variant_flags', [0, 128, 256])
def test_warp_per_problem_structure_skipping_is_bit_identical(T, variant_flags):
    """struct_o = 2 only drops products with exact zeros of the component-major two-output system: same bits as
    struct_o = 0, and the inverse agrees with numpy."""
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(22)
    Ns_list = [202, 201, 200, 150, 37]
    sizes = [2 * ((N + 1) // 2 * 2) for N in Ns_list]
    nmax = max(sizes)
    lda = nmax + 4
    mats = [_struct_spd(rng, N) for N in Ns_list]
    out = []
    lib.kmpx_debug_set(variant_flags)
    try:
        for struct_o in (0, 2):
            A = T.full((len(sizes), nmax, lda), float('nan'), dtype=T.float64, device='cuda')
            for b, n in enumerate(sizes):
                A[b, :n, :n] = dev(T, np.tril(mats[b]) + np.triu(np.full((n, n), np.nan), 1))
            n_sys = dev(T, np.array(sizes, dtype=np.int32))
            info = T.zeros(len(sizes), dtype=T.int32, device='cuda')
            _lib.call('kmpx_potrf_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), _lib.ptr(info), struct_o, _lib.stream())
            _lib.call('kmpx_trtri_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(sizes), struct_o, _lib.stream())
            assert info.cpu().numpy().tolist() == [0] * len(sizes)
            out.append(A.cpu().numpy())
    finally:
        lib.kmpx_debug_set(0)
    for b, n in enumerate(sizes):
        assert np.array_equal(np.tril(out[0][b, :n, :n]), np.tril(out[1][b, :n, :n])), n
        assert rel_err(np.tril(out[1][b, :n, :n]), np.linalg.inv(np.linalg.cholesky(mats[b]))) < 1e-12, n


@pytest.mark.parametrize('variant_flags', [0, 128, 256])
def test_trtri_warp_per_problem_with_fused_solve(T, variant_flags):
    """kmpx_trtri_wpp_solve leaves w = Linv y' next to the inverse (ragged component-major two-output systems, y point-major)."""
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(24)
    pts = [202, 201, 200, 150, 37, 202]
    O = 2
    n_max_pts = max(pts)
    nmax = O * ((n_max_pts + 1) // 2 * 2)
    lda = nmax + 4
    sizes = [O * ((N + 1) // 2 * 2) for N in pts]
    mats = [_struct_spd(rng, N) for N in pts]
    A = T.full((len(pts), nmax, lda), float('nan'), dtype=T.float64, device='cuda')
    for b, n in enumerate(sizes):
        A[b, :n, :n] = dev(T, np.tril(mats[b]) + np.triu(np.full((n, n), np.nan), 1))
    y = rng.normal(size=(len(pts), n_max_pts, O))
    yd = dev(T, y)
    n_sys, n_pts = dev(T, np.array(sizes, dtype=np.int32)), dev(T, np.array(pts, dtype=np.int32))
    info = T.zeros(len(pts), dtype=T.int32, device='cuda')
    w = T.full((len(pts), lda), float('nan'), dtype=T.float64, device='cuda')
    lib.kmpx_debug_set(variant_flags)
    try:
        _lib.call('kmpx_potrf_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(pts), _lib.ptr(info), 2, _lib.stream())
        _lib.call('kmpx_trtri_wpp_solve', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, len(pts), 2,
                  _lib.ptr(yd), _lib.ptr(n_pts), n_max_pts, O, _lib.ptr(w), lda, _lib.stream())
    finally:
        lib.kmpx_debug_set(0)
    assert info.cpu().numpy().tolist() == [0] * len(pts)
    Xh, wh = A.cpu().numpy(), w.cpu().numpy()
    for b, (N, n) in enumerate(zip(pts, sizes)):
        Ns = n // O
        ys = np.zeros(n)
        for a in range(O):
            ys[a * Ns:a * Ns + N] = y[b, :N, a]
        Linv = np.linalg.inv(np.linalg.cholesky(mats[b]))
        assert rel_err(np.tril(Xh[b, :n, :n]), Linv) < 1e-12, N
        assert rel_err(wh[b, :n], Linv @ ys) < 1e-12, N


def test_potrf_warp_per_problem_reports_first_bad_pivot(T):
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(23)
    n = 160
    good = _spd(rng, n)
    bad = good.copy(); bad[100, 100] = -5.0
    bad0 = good.copy(); bad0[0, 0] = 0.0
    A = dev(T, np.stack([good, bad, good, bad0]))
    info = T.zeros(4, dtype=T.int32, device='cuda')
    _lib.call('kmpx_potrf_wpp', _lib.ptr(A), n, n * n, None, n, 4, _lib.ptr(info), 0, _lib.stream())
    assert info.cpu().numpy().tolist() == [0, 101, 0, 1]
    assert rel_err(np.tril(A[2].cpu().numpy()), np.linalg.cholesky(good)) < 1e-13


def test_potrf_reports_first_bad_pivot(T):
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(4)
    n = 70
    S = _spd(rng, n)
    S[40, 40] = -1.0
    A = dev(T, np.stack([S, _spd(rng, n)]))
    info = T.zeros(2, dtype=T.int32, device='cuda')
    _lib.call('kmpx_potrf', _lib.ptr(A), n, n * n, None, n, 2, _lib.ptr(info), None, 0, _lib.stream())
    assert info.cpu().numpy().tolist() == [41, 0]


@pytest.mark.parametrize('n', [600, 1100, 1536])
def test_potrf_trtri_large_blocked(T, n):
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n)
    S = _spd(rng, n)
    lda = (n + 7) // 8 * 8
    A = T.zeros((n, lda), dtype=T.float64, device='cuda')
    A[:, :n] = dev(T, S)
    nbytes = max(int(lib.kmpx_potrf_workspace_bytes(n, 1)), int(lib.kmpx_trtri_workspace_bytes(n, 1)))
    ws = T.empty(nbytes // 8, dtype=T.float64, device='cuda')
    info = T.full((1,), -1, dtype=T.int32, device='cuda')
    _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, _lib.stream())
    assert int(info.item()) == 0
    Lref = np.linalg.cholesky(S)
    assert rel_err(np.tril(A[:, :n].cpu().numpy()), Lref) < 1e-12
    _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(ws), nbytes, _lib.stream())
    assert rel_err(np.tril(A[:, :n].cpu().numpy()), np.linalg.inv(Lref)) < 1e-11


def test_inverse_gauss_jordan(T):
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    sizes = [1, 7, 64, 150, 300]
    nmax = 300
    A = T.zeros((len(sizes), nmax, nmax), dtype=T.float64, device='cuda')
    mats = []
    for b, n in enumerate(sizes):
        M = rng.normal(size=(n, n)) + 0.5 * np.eye(n)
        if n > 1:
            M[0, 0] = 0.0    # forces a row interchange in the first step
        mats.append(M)
        A[b, :n, :n] = dev(T, M)
    nbytes = int(lib.kmpx_inverse_workspace_bytes(nmax, len(sizes)))
    ws = T.empty(nbytes // 8 + 1, dtype=T.float64, device='cuda')
    info = T.zeros(len(sizes), dtype=T.int32, device='cuda')
    n_d = dev(T, np.array(sizes, dtype=np.int32))
    _lib.call('kmpx_inverse', _lib.ptr(A), nmax, nmax * nmax, _lib.ptr(n_d), nmax, len(sizes),
              _lib.ptr(info), _lib.ptr(ws), nbytes, _lib.stream())
    out = A.cpu().numpy()
    assert np.all(info.cpu().numpy() == 0)
    for b, n in enumerate(sizes):
        ref = np.linalg.inv(mats[b])
        assert rel_err(out[b, :n, :n], ref) < 1e-9 * np.linalg.cond(mats[b]) / 1e3 + 1e-11, n
    sing = T.zeros((1, 4, 4), dtype=T.float64, device='cuda')
    _lib.call('kmpx_inverse', _lib.ptr(sing), 4, 16, None, 4, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, _lib.stream())
    assert int(info[0].item()) > 0


def test_branch_free_exponentials_are_accurate_to_an_ulp_or_two(T):
    """exp_nonpos / exp_tab (common.cuh) against the correctly rounded exp of mpmath-grade numpy long doubles."""
    from kmp_demos_b200 import _lib
    rng = np.random.default_rng(3)
    x = np.concatenate([-rng.random(200000) * 50.0, -10.0 ** rng.uniform(-12, 2.8, 100000), rng.random(50000) * 20.0,
                        [0.0, -0.0, -708.0, -707.9, -1e-320, -745.0, -1e4, -np.inf, np.nan]])
    xd = T.as_tensor(x, device='cuda')
    ref = np.exp(x.astype(np.longdouble))
    for mode in (0, 1):
        y = T.empty_like(xd)
        _lib.call('kmpx_debug_exp', _lib.ptr(xd), _lib.ptr(y), x.size, mode, _lib.stream())
        y = y.cpu().numpy()
        ok = np.isfinite(x) & (x >= -708.0)
        ulp = np.spacing(np.exp(x[ok]))
        err = np.abs(y[ok].astype(np.longdouble) - ref[ok]) / ulp
        assert float(err.max()) < 2.0, (mode, float(err.max()))
        assert np.all(y[np.isfinite(x) & (x < -708.0)] == 0.0) and y[-2] == 0.0 and np.isnan(y[-1])


def test_gmr_matches_reference(T, golden):
    from kmp_demos_b200.mixture import GaussianMixtureModel
    z = golden('small_cases.npz')
    for c in range(int(z['gmr_cases'])):
        I, O, K, M, reg = z[f'gmr{c}_cfg']
        g = GaussianMixtureModel(n_components=int(K), diag_reg_factor=float(reg))
        g.n_features = int(I + O)
        g.priors, g.means, g.covariances = z[f'gmr{c}_priors'], z[f'gmr{c}_means'], z[f'gmr{c}_covs']
        mu, sg = g.predict(z[f'gmr{c}_x'])
        assert rel_err(mu, z[f'gmr{c}_mu']) < 1e-12 and rel_err(sg, z[f'gmr{c}_sigma']) < 1e-11, c
        assert np.all(mu[:, -1] == 0) and rel_err(sg[:, :, -1], float(reg) * np.eye(int(O))) < 1e-15   # all-underflow query


def test_gmr_time_grid_with_component_skipping_vs_oracle(T, golden):
    """The scalar-input GMR path skips components more than 60 ln 2 below the leading exponent (k_gmr): on a dense time
    grid through the reference's own cfg-1 mixtures - where most components are skipped - the result must equal the
    oracle's full sum to 1e-12, far-away and NaN queries included."""
    from kmp_demos_b200.mixture import GaussianMixtureModel
    from oracle import kmp_oracle as orc
    z = golden('letters_ref.npz')
    x = np.concatenate([np.linspace(-0.5, 2.5, 20001), [50.0, -1e3]]).reshape(1, -1)
    for li in (0, 6, 18, 25):
        g = GaussianMixtureModel(n_components=8, diag_reg_factor=1e-6)
        g.n_features = 3
        g.priors, g.means, g.covariances = z['priors'][li], z['means'][li], z['covariances'][li]
        mu, sg = g.predict(x)
        mr, sr = orc.gmr(g.priors, g.means, g.covariances, x, 1e-6)
        assert np.max(np.abs(mu - mr)) < 1e-12 * max(1.0, np.max(np.abs(mr))), li
        assert np.max(np.abs(sg - sr)) < 1e-12 * max(1.0, np.max(np.abs(sr))), li
        assert np.array_equal(sg[0, 1], sg[1, 0])
    xn = np.array([[0.5, np.nan, 1.0]])
    mu, sg = g.predict(xn)
    assert np.isnan(mu[:, 1]).all() and np.isfinite(mu[:, [0, 2]]).all()


def test_reference_kmeans_labels_bit_exact(T, golden):
    from kmp_demos_b200.cluster import KMeans, Uniform
    z = golden('small_cases.npz')
    for c in range(int(z['km_cases'])):
        np.random.seed(int(z[f'km{c}_seed']))
        labels = KMeans(int(z[f'km{c}_K'])).fit_predict(z[f'km{c}_X'])
        assert labels.dtype == np.int64 and np.array_equal(labels, z[f'km{c}_labels']), c
    assert np.array_equal(Uniform(8, 5).fit_predict(np.zeros((3, 1000))), z['uniform_1000_8_5'])
    with pytest.raises(ValueError):
        KMeans(0)


def test_reference_kmeans_empty_cluster_nan_rule(T):
    """An empty cluster gives a NaN centroid and numpy's argmin then sends every sample to it (Cluster.py:91-95)."""
    from kmp_demos_b200.cluster import KMeans
    from oracle import kmp_oracle as orc
    X = np.array([[0.0, 0.0, 0.0, 10.0, 10.0, 10.0], [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]])   # duplicates => empty cluster
    np.random.seed(0)
    init = np.random.choice(6, 3, replace=False)
    np.random.seed(0)
    got = KMeans(3, n_iter=5).fit_predict(X)
    ref, _ = orc.ref_kmeans(X, 3, 5, 1e-4, init)
    assert np.array_equal(got, ref)


def test_quaternion_maps(T, golden):
    import kmp_demos_b200.types.quaternion as qm
    from kmp_demos_b200.types.quaternion import quaternion
    z = golden('small_cases.npz')
    Q, P, W = z['quat_Q'], z['quat_P'], z['quat_W']
    qn = qm.normalize_batch(Q)
    pn = qm.normalize_batch(P)
    assert rel_err(qn, z['quat_norm']) < 1e-15
    assert np.max(np.abs(qm.log_batch(qn) - z['quat_log'])) < 1e-14
    assert np.max(np.abs(qm.exp_batch(W) - z['quat_exp'])) < 1e-15
    assert np.max(np.abs(qm.mul_batch(qn, pn) - z['quat_mul'])) < 1e-15
    assert np.max(np.abs(qm.log_batch(qm.mul_batch(qn, pn, conj_q=True)) - z['quat_mulconj_log'])) < 1e-13
    rv = qm.from_rotvec_batch(W)
    assert np.allclose(rv, z['quat_rotvec'], rtol=0, atol=1e-15, equal_nan=True) and np.all(np.isnan(rv[0]))
    # scalar class, SURVEY.md 8c known answers
    q = quaternion(-0.2, np.array([0.8, -0.7, -1.6]))
    assert np.allclose(q.as_array(), [-0.10355607461569954, 0.4142242984627982, -0.3624462611549484, -0.8284485969255964], atol=1e-16)
    assert np.allclose(q.log(), [-0.610974356491441, 0.5346025619300109, 1.221948712982882], atol=1e-15)
    assert np.array_equal(quaternion.exp(np.array([1e-9] * 3)).as_array(), [1, 0, 0, 0])
    assert np.allclose(quaternion.exp(np.array([2e-8, 0, 0])).as_array(), [0.9999999999999998, 2e-8, 0, 0], atol=1e-16)
    assert np.array_equal(quaternion(0, np.array([1.0, 0, 0])).log(), [0, 0, 0])
    r = quaternion(1, np.array([2., 3, 4])) * quaternion(.5, np.array([-1, .2, .3]))
    assert np.allclose(r.as_array(), [0.10879222762803663, 0.01554174680400521, -0.45071065731615173, 0.8858795678282981], atol=1e-15)
    w = np.array([0.508369, 0.857956, -0.167863])
    assert np.allclose(quaternion.exp(w).log(), w, atol=1e-15)
    assert np.allclose((~q).as_array(), q.as_array() * [1, -1, -1, -1], atol=1e-16)
    assert np.allclose(quaternion.from_rotation_vector(np.array([3., 2, 1])).as_array(),
                       [0.29555112749297824, 0.7659655801357929, 0.5106437200905286, 0.2553218600452643], atol=1e-15)


def test_em_statistics_partition_sum_equals_whole(T):
    """Data partitioned in two: the summed statistics (what the NCCL all-reduce produces) give the same M-step
    as the unpartitioned pass (<= 1e-12)."""
    from kmp_demos_b200.mixture import EMEngine
    from kmp_demos_b200 import workloads as wl
    from oracle import kmp_oracle as orc
    X, means, _ = wl.cfg5_samples(30000, K=12, d=7, seed=1)
    lab = np.argmin(((X[:, None, :] - means[None]) ** 2).sum(-1), axis=1).astype(np.int32)
    Xd = dev(T, X)
    shift = dev(T, X.mean(axis=0))
    whole = EMEngine(Xd, 12, 1e-6, shift=shift)
    whole.init_from_labels(dev(T, lab))
    # oracle M-step from the same labels
    resp = np.zeros((X.shape[0], 12)); resp[np.arange(X.shape[0]), lab] = 1
    nk, mu, cov = orc.gmm_m_step(X, resp, 1e-6)
    assert rel_err(whole.means.cpu().numpy(), mu) < 1e-12 and rel_err(whole.covs.cpu().numpy(), cov) < 1e-10
    assert rel_err(whole.prec_chol.cpu().numpy(), orc.gmm_precisions_chol(cov)) < 1e-9
    whole.em_pass()
    lse, log_resp = orc.gmm_e_step(X, nk / X.shape[0], mu, orc.gmm_precisions_chol(cov))
    assert abs(float(whole.stats[-1].item()) - lse.sum()) < 1e-9 * abs(lse.sum())
    nk2, mu2, cov2 = orc.gmm_m_step(X, np.exp(log_resp), 1e-6)
    assert rel_err(whole.means.cpu().numpy(), mu2) < 1e-11 and rel_err(whole.covs.cpu().numpy(), cov2) < 1e-9
    half = X.shape[0] // 2 + 37
    parts = [EMEngine(dev(T, X[:half]), 12, 1e-6, shift=shift, n_total=X.shape[0]),
             EMEngine(dev(T, X[half:]), 12, 1e-6, shift=shift, n_total=X.shape[0])]
    labs = [dev(T, lab[:half]), dev(T, lab[half:])]
    from kmp_demos_b200 import _lib
    for p, l in zip(parts, labs):
        _lib.call('kmpx_gmm_label_stats', _lib.ptr(p.X), p.n, p.d, p.K, _lib.ptr(p.shift), _lib.ptr(l), _lib.ptr(p.stats),
                  _lib.ptr(p.ws), p.ws_bytes, _lib.stream())
    tot = parts[0].stats + parts[1].stats
    for p in parts:
        p.stats.copy_(tot); p._finalize(first=True)
    for p in parts:
        _lib.call('kmpx_gmm_em_pass', _lib.ptr(p.X), p.n, p.d, p.K, _lib.ptr(p.shift), _lib.ptr(p.log_weights), _lib.ptr(p.means),
                  _lib.ptr(p.prec_chol), _lib.ptr(p.log_det), _lib.ptr(p.stats), None, None, _lib.ptr(p.ws), p.ws_bytes, _lib.stream())
    tot = parts[0].stats + parts[1].stats
    parts[0].stats.copy_(tot); parts[0]._finalize(first=False)
    assert rel_err(parts[0].means.cpu().numpy(), whole.means.cpu().numpy()) < 1e-12
    assert rel_err(parts[0].covs.cpu().numpy(), whole.covs.cpu().numpy()) < 1e-12
