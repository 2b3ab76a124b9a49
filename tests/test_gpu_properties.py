"""Property-based differential tests: hypothesis draws shapes, hyper-parameters and
seeds; the drop-in classes on the GPU are compared with the CPU oracle (oracle/kmp_oracle.py, pinned against the live
reference in tests/test_oracle_golden.py) on the same inputs.  Tolerances are the north star's: <= 1e-9 relative on
means, <= 1e-7 on covariances (norm-wise, scaled by the conditioning of the drawn system where it exceeds the demos'),
bit-exact for labels and indices."""
import warnings

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from conftest import rel_err
from kmp_demos_b200 import workloads as wl  # noqa: F401
from oracle import kmp_oracle as orc

pytestmark = pytest.mark.gpu
COMMON = dict(deadline=None, derandomize=True, suppress_health_check=list(HealthCheck), print_blob=True)


def _spd_blocks(rng, n, O, lo=0.05):
    B = rng.normal(size=(n, O, O))
    return np.transpose(np.einsum('nab,ncb->nac', B, B) + lo * np.eye(O), (1, 2, 0)).copy()


@settings(max_examples=25, **COMMON)
@given(I=st.integers(1, 2), O=st.integers(1, 6), N=st.integers(2, 45), M=st.integers(1, 17),
       sigma_f=st.floats(0.5, 300.0), lam=st.floats(0.05, 3.0), alpha=st.floats(0.5, 80.0), td=st.booleans(),
       seed=st.integers(0, 2 ** 31 - 1))
def test_kmp_fit_predict_matches_oracle(I, O, N, M, sigma_f, lam, alpha, td, seed):
    """KMP.fit / predict (kernel build, Cholesky, triangular inverse, fused predict; the generic kernel for O > 2 and the
    time-driven kernel) == the restatement of KMP.py:130-194."""
    from kmp.model import KMP
    rng = np.random.default_rng(seed)
    td = td and O % 2 == 0
    s = np.sort(rng.random((I, N)), axis=1) * 2.0
    mu = rng.normal(size=(O, N)) * 3.0
    sig = _spd_blocks(rng, N, O)
    q = rng.random((I, M)) * 2.0
    A = orc.kmp_matrix(s, sig, lam, sigma_f, td)
    cond = np.linalg.cond(A)
    if not np.isfinite(cond) or cond > 1e10:
        return                                           # numerically singular draw: no reference to compare with
    est = np.linalg.inv(A)
    xr, sr = orc.kmp_predict(est, s, mu, q, alpha, sigma_f, td)
    k = KMP(l=lam, alpha=alpha, sigma_f=sigma_f, time_driven_kernel=td)
    k._logger.disabled = True
    k.fit(s, mu, sig)
    xi, sg = k.predict(q)
    slack = max(1.0, cond / 1e5)                         # the demos' systems have cond <= 4e6 at the stated tolerances
    tol_m, tol_c = (1e-7, 1e-7) if td else (1e-9, 1e-7)
    assert xi.shape == (O, M) and sg.shape == (O, O, M)
    assert rel_err(xi, xr) < tol_m * slack, (rel_err(xi, xr), cond)
    assert np.max(np.abs(sg - sr)) < tol_c * slack * max(np.max(np.abs(sr)), alpha), (np.max(np.abs(sg - sr)), cond)


@settings(max_examples=12, **COMMON)
@given(I=st.integers(1, 2), O=st.integers(1, 2), N=st.integers(2, 230), M=st.integers(1, 150),
       sigma_f=st.sampled_from([0.5, 5.0, 50.0, 200.0, 1000.0]), seed=st.integers(0, 2 ** 31 - 1))
def test_fused_predict_both_cta_shapes_match_oracle(I, O, N, M, sigma_f, seed):
    """The table-driven fused predict kernel (plain kernel, O <= 2) in both CTA shapes - 32 queries with two CTAs per SM and
    64 queries with one (kmpx_predict_tile_queries) - with the kappa-box skipping on, on queries that reach outside the
    data: each == the restatement of KMP.py:160-194, and the two shapes agree to rounding."""
    from kmp.model import KMP
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(seed)
    s = np.sort(rng.random((I, N)), axis=1) * 2.0
    mu = rng.normal(size=(O, N)) * 3.0
    sig = _spd_blocks(rng, N, O)
    q = rng.random((I, M)) * 4.0 - 1.0
    A = orc.kmp_matrix(s, sig, 0.5, sigma_f, False)
    cond = np.linalg.cond(A)
    if not np.isfinite(cond) or cond > 1e10:
        return
    xr, sr = orc.kmp_predict(np.linalg.inv(A), s, mu, q, 40.0, sigma_f, False)
    k = KMP(l=0.5, alpha=40.0, sigma_f=sigma_f, time_driven_kernel=False)
    k._logger.disabled = True
    k.fit(s, mu, sig)
    out = {}
    try:
        for tile in (32, 64):
            assert lib.kmpx_predict_tile_queries(tile) == 0
            out[tile] = k.predict(q)
    finally:
        assert lib.kmpx_predict_tile_queries(0) == 0
    slack = max(1.0, cond / 1e5)
    for tile, (xi, sg) in out.items():
        assert np.max(np.abs(xi - xr)) < 1e-9 * slack * max(np.max(np.abs(xr)), 1.0), (tile, np.max(np.abs(xi - xr)), cond)
        assert np.max(np.abs(sg - sr)) < 1e-7 * slack * 40.0, (tile, np.max(np.abs(sg - sr)), cond)
    assert np.max(np.abs(out[32][0] - out[64][0])) < 1e-11 * slack * max(np.max(np.abs(xr)), 1.0)
    assert np.max(np.abs(out[32][1] - out[64][1])) < 1e-11 * slack * 40.0


@settings(max_examples=20, **COMMON)
@given(O=st.integers(1, 3), N=st.integers(3, 30), P=st.integers(1, 4), seed=st.integers(0, 2 ** 31 - 1),
       near=st.lists(st.booleans(), min_size=4, max_size=4))
def test_set_waypoint_replace_or_append_matches_oracle(O, N, P, seed, near):
    """KMP.set_waypoint (nearest stored input, strict < tol, first minimum, replace in place or append, refit)
    == KMP.py:57-91: same stored arrays, same predictions."""
    from kmp.model import KMP
    rng = np.random.default_rng(seed)
    s = (0.01 * np.arange(1, N + 1)).reshape(1, -1)
    mu = rng.normal(size=(O, N))
    sig = _spd_blocks(rng, N, O)
    vs, vx, vg = [], [], []
    for j in range(P):
        if near[j]:
            base = s[0, rng.integers(0, N)]
            vs.append(float(base + rng.choice([0.0, 3e-4, -4.9e-4, 5.1e-4])))       # inside and just outside tol = 5e-4
        else:
            vs.append(float(rng.uniform(0.0, 0.012 * N)))
        vx.append(rng.normal(size=(1, O)))
        vg.append(1e-4 * np.eye(O) * (1 + rng.random()))
    s2, y2, g2 = orc.kmp_set_waypoint(s, mu, sig, vs, vx, vg, 5e-4)
    k = KMP(l=0.5, alpha=10.0, sigma_f=200.0, time_driven_kernel=False)
    k._logger.disabled = True
    k.fit(s, mu, sig)
    k.set_waypoint(vs, vx, vg)
    assert k.N == s2.shape[1] and np.array_equal(k.s, s2) and np.array_equal(k.xi, y2) and np.array_equal(k.sigma, g2)
    q = np.linspace(0.0, 0.012 * N, 23).reshape(1, -1)
    est = orc.kmp_fit(s2, g2, 0.5, 200.0, False)
    xr, sr = orc.kmp_predict(est, s2, y2, q, 10.0, 200.0, False)
    xi, sg = k.predict(q)
    assert rel_err(xi, xr) < 1e-9 and np.max(np.abs(sg - sr)) < 1e-7 * max(np.max(np.abs(sr)), 10.0)


@settings(max_examples=25, **COMMON)
@given(I=st.integers(1, 3), O=st.integers(1, 6), K=st.integers(1, 14), M=st.integers(1, 70), reg=st.floats(1e-8, 1e-2),
       spread=st.floats(0.2, 5.0), seed=st.integers(0, 2 ** 31 - 1))
def test_gmr_matches_oracle(I, O, K, M, reg, spread, seed):
    """GaussianMixtureModel.predict (k_gmr: scalar-input path with component skipping for I = 1, generic path otherwise)
    == the closed form of GaussianMixtureModel.py:83-105."""
    from kmp.mixture import GaussianMixtureModel
    rng = np.random.default_rng(seed)
    d = I + O
    g = GaussianMixtureModel(n_components=K, diag_reg_factor=reg)
    g.logger.disabled = True
    g.n_features = d
    pri = rng.random(K) + 0.05
    g.priors = pri / pri.sum()
    g.means = rng.normal(size=(d, K)) * spread
    B = rng.normal(size=(K, d, d)) * 0.5
    g.covariances = np.transpose(np.einsum('kab,kcb->kac', B, B) + 0.05 * np.eye(d), (1, 2, 0)).copy()
    x = rng.normal(size=(I, M)) * spread
    mu, sg = g.predict(x)
    mr, sr = orc.gmr(g.priors, g.means, g.covariances, x, reg)
    assert np.max(np.abs(mu - mr)) < 1e-11 * max(1.0, np.max(np.abs(mr)))
    assert np.max(np.abs(sg - sr)) < 1e-10 * max(1.0, np.max(np.abs(sr)))


@settings(max_examples=20, **COMMON)
@given(d=st.integers(1, 9), n=st.integers(8, 400), K=st.integers(1, 8), seed=st.integers(0, 2 ** 31 - 1),
       dup=st.booleans())
def test_reference_kmeans_labels_bit_exact(d, n, K, seed, dup):
    """kmp.cluster.KMeans (Cluster.py:74-102) through the bit-exact kernels == the literal restatement: int64 labels."""
    from kmp.cluster import KMeans
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(d, n)) + 2.5 * rng.integers(0, 3, size=(1, n))
    if dup:
        X[:, n // 2:] = X[:, :n - n // 2]                # duplicated samples: exact distance ties
    K = min(K, n)
    np.random.seed(seed % (2 ** 32))
    init = np.random.choice(n, K, replace=False)
    np.random.seed(seed % (2 ** 32))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        got = KMeans(K, n_iter=25).fit_predict(X)
        ref, _ = orc.ref_kmeans(X, K, 25, 1e-4, init)
    assert got.dtype == np.int64 and np.array_equal(got, ref)


@settings(max_examples=30, **COMMON)
@given(seed=st.integers(0, 2 ** 31 - 1), scale=st.sampled_from([1e-9, 1e-7, 1e-3, 0.3, 1.0, 3.0]))
def test_quaternion_maps_match_oracle_and_round_trip(seed, scale):
    """Batched log / exp / product (quaternion.py:60-96,182-202) == the restatement, and exp(log(q)) = +-q."""
    import kmp_demos_b200.types.quaternion as qm
    rng = np.random.default_rng(seed)
    Q = rng.normal(size=(33, 4)); P = rng.normal(size=(33, 4)); W = rng.normal(size=(33, 3)) * scale
    qn = qm.normalize_batch(Q)
    assert np.max(np.abs(qn - np.stack([orc.quat_normalize(q) for q in Q]))) < 1e-15
    lg = qm.log_batch(qn)
    assert np.max(np.abs(lg - np.stack([orc.quat_log(q) for q in qn]))) < 1e-13
    ex = qm.exp_batch(W)
    assert np.max(np.abs(ex - np.stack([orc.quat_exp(w) for w in W]))) < 1e-15
    back = qm.exp_batch(lg)
    sign = np.sign(np.sum(back * qn, axis=1, keepdims=True))
    assert np.max(np.abs(back * sign - qn)) < 1e-13
    pn = qm.normalize_batch(P)
    assert np.max(np.abs(qm.mul_batch(qn, pn) - np.stack([orc.quat_mul(a, b) for a, b in zip(qn, pn)]))) < 1e-15


@settings(max_examples=8, **COMMON)
@given(n=st.integers(300, 4000), d=st.integers(1, 7), K=st.integers(1, 9), seed=st.integers(0, 2 ** 31 - 1))
def test_gmm_fit_matches_sklearn_restatement(n, d, K, seed):
    """GaussianMixtureModel.fit (device k-means++ / Lloyd / EM) == the numpy restatement of sklearn's
    GaussianMixture(random_state=420): iteration count, lower bound, parameters, argmax labels."""
    from kmp.mixture import GaussianMixtureModel
    X, _, _ = wl.cfg5_samples(n, K=max(K, 2), d=7, seed=seed % 1000)
    X = np.ascontiguousarray(X[:, :d])
    r = orc.gmm_fit(X, K, 1e-6)
    g = GaussianMixtureModel(n_components=K, diag_reg_factor=1e-6)
    g.logger.disabled = True
    g.fit(X)
    assert g.model.n_iter_ == r['n_iter']
    assert abs(g.model.lower_bound_ - r['lower_bound']) < 1e-9 * max(1.0, abs(r['lower_bound']))
    assert rel_err(g.priors, r['weights']) < 1e-9 and rel_err(g.means, r['means'].T) < 1e-9
    assert rel_err(g.covariances, np.transpose(r['covariances'], (1, 2, 0))) < 1e-7
