"""Pins the CPU oracle (oracle/kmp_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from oracle import kmp_oracle as orc
from kmp_demos_b200 import workloads as wl
from conftest import rel_err

MEAN_TOL = 1e-9     # north-star tolerance on means
COV_TOL = 1e-7      # north-star tolerance on covariances


def test_kernel_matrix_and_predict_small(golden):
    z = golden('small_cases.npz')
    for c in range(int(z['kmp_cases'])):
        I, O, N, td, sf, lam, alpha = z[f'kmp{c}_cfg']
        O, N, td = int(O), int(N), bool(td)
        s, mu, sig = z[f'kmp{c}_s'], z[f'kmp{c}_mu'], z[f'kmp{c}_sigma']
        A_lit = orc.kmp_matrix_literal(s, sig, lam, sf, td)
        A_vec = orc.kmp_matrix(s, sig, lam, sf, td)
        assert np.array_equal(A_lit, z[f'kmp{c}_A']), f'case {c}: literal matrix is not bit-exact'
        # H3: the finite-difference blocks divide 1-ulp exp differences (scalar vs SIMD exp) by dt^2=1e-6
        assert rel_err(A_vec, z[f'kmp{c}_A']) < (1e-11 if td else 1e-15)
        est = orc.kmp_fit(s, sig, lam, sf, td, literal=True)
        assert np.array_equal(est, z[f'kmp{c}_estimator'])
        xi, sg = orc.kmp_predict_literal(est, s, mu, z[f'kmp{c}_q'], alpha, sf, td)
        assert np.array_equal(xi, z[f'kmp{c}_xi']) and np.array_equal(sg, z[f'kmp{c}_sg'])
        xi2, sg2 = orc.kmp_predict(est, s, mu, z[f'kmp{c}_q'], alpha, sf, td)
        assert rel_err(xi2, xi) < (1e-8 if td else 1e-11) and rel_err(sg2, sg) < (1e-7 if td else 1e-9)
        assert np.array_equal(orc.kernel_block(z[f'kmp{c}_q'][:, 0], z[f'kmp{c}_q'][:, 0], sf, O, td), z[f'kmp{c}_kss'])


def test_gmr_small(golden):
    z = golden('small_cases.npz')
    for c in range(int(z['gmr_cases'])):
        reg = float(z[f'gmr{c}_cfg'][4])
        args = (z[f'gmr{c}_priors'], z[f'gmr{c}_means'], z[f'gmr{c}_covs'], z[f'gmr{c}_x'], reg)
        mu, sg = orc.gmr_literal(*args)
        assert rel_err(mu, z[f'gmr{c}_mu']) < 1e-13 and rel_err(sg, z[f'gmr{c}_sigma']) < 1e-12
        mu2, sg2 = orc.gmr(*args)
        assert rel_err(mu2, z[f'gmr{c}_mu']) < 1e-12 and rel_err(sg2, z[f'gmr{c}_sigma']) < 1e-11
        # far-away query: zero activation -> mean 0, covariance reg*I (GMM.py:90 quirk)
        assert np.all(mu2[:, -1] == 0) and np.allclose(sg2[:, :, -1], reg * np.eye(mu2.shape[0]), atol=0, rtol=1e-15)


def test_ref_kmeans_and_uniform(golden):
    z = golden('small_cases.npz')
    for c in range(int(z['km_cases'])):
        labels, _ = orc.ref_kmeans(z[f'km{c}_X'], int(z[f'km{c}_K']), 100, 1e-4, z[f'km{c}_init'])
        assert np.array_equal(labels, z[f'km{c}_labels'])
    assert np.array_equal(orc.uniform_labels(1000, 8, 5), z['uniform_1000_8_5'])
    assert np.array_equal(orc.uniform_labels(903, 7, 3), z['uniform_903_7_3'])


def test_quaternion_maps(golden):
    z = golden('small_cases.npz')
    Q, P, W = z['quat_Q'], z['quat_P'], z['quat_W']
    for i in range(Q.shape[0]):
        q = orc.quat_normalize(Q[i])
        p = orc.quat_normalize(P[i])
        assert np.allclose(q, z['quat_norm'][i], rtol=0, atol=1e-16)
        assert np.allclose(orc.quat_log(q), z['quat_log'][i], rtol=0, atol=2e-16)
        assert np.allclose(orc.quat_mul(q, p), z['quat_mul'][i], rtol=0, atol=2e-16)
        assert np.allclose(orc.quat_conj(q), z['quat_conj'][i], rtol=0, atol=2e-16)
        assert np.allclose(orc.quat_exp(W[i]), z['quat_exp'][i], rtol=0, atol=2e-16)
        assert np.allclose(orc.quat_log(orc.quat_mul(q, orc.quat_conj(p))), z['quat_mulconj_log'][i], rtol=0, atol=1e-15)
        assert np.allclose(orc.quat_from_rotation_vector(W[i]), z['quat_rotvec'][i], rtol=0, atol=2e-16, equal_nan=True)
    # SURVEY.md 8c known answers
    assert np.allclose(orc.quat_normalize([-0.2, 0.8, -0.7, -1.6]),
                       [-0.10355607461569954, 0.4142242984627982, -0.3624462611549484, -0.8284485969255964], atol=1e-16)
    assert np.array_equal(orc.quat_exp([1e-9] * 3), [1, 0, 0, 0])
    assert np.array_equal(orc.quat_log([0, 1, 0, 0]), [0, 0, 0])


@pytest.mark.parametrize('letter', ['A', 'G', 'S', 'Z'])
def test_gmm_fit_letters(golden, letter):
    """sklearn restatement (k-means++ / Lloyd / EM) vs the reference's GaussianMixtureModel.fit."""
    z = golden('letters_ref.npz')
    li = wl.LETTERS.index(letter)
    X = wl.cfg1_design_matrix(z['pos'][li])
    r = orc.gmm_fit(X, 8, 1e-6)
    assert r['n_iter'] == int(z['n_iter'][li])
    assert abs(r['lower_bound'] - float(z['lower_bound'][li])) < 1e-11
    assert rel_err(r['weights'], z['priors'][li]) < 1e-10
    assert rel_err(r['means'].T, z['means'][li]) < 1e-10
    assert rel_err(np.transpose(r['covariances'], (1, 2, 0)), z['covariances'][li]) < 1e-9
    mu, sg = orc.gmr(r['weights'], r['means'].T, np.transpose(r['covariances'], (1, 2, 0)), wl.cfg1_gmr_grid(200), 1e-6)
    assert rel_err(mu, z['gmr_mu'][li]) < MEAN_TOL and rel_err(sg, z['gmr_sigma'][li]) < COV_TOL


def test_gmm_fit_synthetic_7d(golden):
    z = golden('gmm_synth.npz')
    X, _, _ = wl.cfg5_samples(int(z['n']), K=int(z['K']), d=7, seed=int(z['seed']))
    r = orc.gmm_fit(X, int(z['K']), 1e-6)
    assert r['n_iter'] == int(z['n_iter'])
    assert rel_err(r['means'].T, z['means']) < 1e-10 and rel_err(r['weights'], z['priors']) < 1e-10
    assert abs(orc.gmm_bic(X, r['weights'], r['means'], r['covariances']) - float(z['bic'])) < 1e-6 * abs(float(z['bic']))


def test_gmm_fit_cfg5_shape_k64(golden):
    """The oracle at cfg 5's own K (64 components, 200 k rows; ~75 s on 8 cores): same iteration count and lower bound
    as sklearn through the reference wrapper, parameters to 1e-10."""
    z = golden('gmm_cfg5_k64.npz')
    X, _, _ = wl.cfg5_samples(int(z['n']), K=int(z['K']), d=7, seed=int(z['seed']))
    r = orc.gmm_fit(X, int(z['K']), 1e-6)
    assert r['n_iter'] == int(z['n_iter'])
    assert abs(r['lower_bound'] - float(z['lower_bound'])) < 1e-11
    assert rel_err(r['means'].T, z['means']) < 1e-10 and rel_err(r['weights'], z['priors']) < 1e-10
    assert rel_err(np.transpose(r['covariances'], (1, 2, 0)), z['covs']) < 1e-9


def test_cfg1_kmp_pipeline(golden):
    zl = golden('letters_ref.npz')
    z = golden('cfg1_G.npz')
    gi = wl.LETTERS.index('G')
    mu, sig = zl['gmr_mu'][gi], zl['gmr_sigma'][gi]
    c = wl.CFG1
    t = wl.cfg1_kmp_grid(200)
    q = wl.cfg1_query_grid(1000)
    est = orc.kmp_fit(t, sig, c['lam'], c['sigma_f'], False)
    xi, sg = orc.kmp_predict(est, t, mu, q, c['alpha'], c['sigma_f'], False)
    assert rel_err(xi, z['n200_pred_xi']) < MEAN_TOL and rel_err(sg, z['n200_pred_sigma']) < COV_TOL
    s2, xi2, sig2 = orc.kmp_set_waypoint(t, mu, sig, list(z['n200_via_t']), [p for p in z['n200_via_p']],
                                         [v for v in z['n200_via_var']], 5e-4)
    assert s2.shape[1] == int(z['n200_adapt_N']) == 201 and np.array_equal(s2, z['n200_adapt_s'])
    est = orc.kmp_fit(s2, sig2, c['lam'], c['sigma_f'], False)
    xi, sg = orc.kmp_predict(est, s2, xi2, q, c['alpha'], c['sigma_f'], False)
    assert rel_err(xi, z['n200_adapt_xi']) < MEAN_TOL and rel_err(sg, z['n200_adapt_sigma']) < COV_TOL


def test_demo1_literal_bit_exact(golden):
    z = golden('cfg1_G.npz')
    T = z['d1_gmr_mu'].shape[1]
    tk = 0.01 * np.arange(1, T + 1).reshape(1, -1)
    est = orc.kmp_fit(tk, z['d1_gmr_sigma'], 0.5, 200.0, False, literal=True)
    assert np.array_equal(est, z['d1_estimator'])
    xi, sg = orc.kmp_predict_literal(est, tk, z['d1_gmr_mu'], tk, 40.0, 200.0, False)
    assert np.array_equal(xi, z['d1_pred_xi']) and np.array_equal(sg, z['d1_pred_sigma'])
    r = orc.gmm_fit(z['d1_X'], 8, 1e-6)
    assert r['n_iter'] == int(z['d1_n_iter']) == 33
    assert rel_err(r['means'].T, z['d1_means']) < 1e-10


def test_cfg3_sample(golden):
    zl = golden('letters_ref.npz')
    z = golden('cfg3_sample.npz')
    c = wl.CFG1
    t = wl.cfg1_kmp_grid(200)
    for r in range(0, len(z['letter']), 5):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        mu, sig = zl['gmr_mu'][li], zl['gmr_sigma'][li]
        vt, vp, vv = wl.cfg3_via_sets(li, mu, si + 1)
        s2, xi2, sig2 = orc.kmp_set_waypoint(t, mu, sig, list(vt[si]), [vp[si, j].reshape(1, -1) for j in range(2)],
                                             [vv[si, j] for j in range(2)], 5e-4)
        assert s2.shape[1] == int(z['N'][r])
        est = orc.kmp_fit(s2, sig2, c['lam'], c['sigma_f'], False)
        xi, sg = orc.kmp_predict(est, s2, xi2, z['queries'], c['alpha'], c['sigma_f'], False)
        assert rel_err(xi, z['xi'][r]) < MEAN_TOL and rel_err(sg, z['sigma'][r]) < COV_TOL


def test_demo2_quirk_and_demo3(golden):
    z = golden('demo2_pose.npz')
    ts = z['time'][:, :150]
    for name in ('pos', 'quat'):
        mu, sig = z[f'{name}_gmr_mu'], z[f'{name}_gmr_sigma']
        # un-listed sigma => broadcast row block (KMP.py:84)
        s2, xi2, sig2 = orc.kmp_set_waypoint(ts, mu, sig, [1], z[f'{name}_wp'].reshape(1, -1), np.eye(3) * 1e-6, 5e-4)
        assert np.array_equal(sig2, z[f'{name}_adapt_sigma_db'])
        est = orc.kmp_fit(s2, sig2, 0.5, 4.0, False)
        xi, sg = orc.kmp_predict(est, s2, xi2, ts, 60.0, 4.0, False)
        assert rel_err(xi, z[f'{name}_adapt_xi']) < MEAN_TOL and rel_err(sg, z[f'{name}_adapt_sigma']) < COV_TOL
    back = np.stack([orc.quat_mul(orc.quat_exp(z['quat_adapt_xi'][:3, i]), z['qa']) for i in range(150)], axis=1)
    assert rel_err(back, z['quat_adapt_back']) < 1e-9
    z3 = golden('demo3_vel.npz')
    tk = 0.01 * np.arange(1, 201).reshape(1, -1)
    s2, xi2, sig2 = orc.kmp_set_waypoint(tk, z3['gmr_mu'], z3['gmr_sigma'], list(z3['via_t']), [p for p in z3['via_p']],
                                         [np.eye(4) * 1e-6] * 4, 5e-4)
    assert s2.shape[1] == int(z3['adapt_N'])
    est = orc.kmp_fit(s2, sig2, 1.0, 6.0, True)
    xi, sg = orc.kmp_predict(est, s2, xi2, tk, 40.0, 6.0, True)
    assert rel_err(xi, z3['adapt_xi']) < 1e-7 and rel_err(sg, z3['adapt_sigma']) < 1e-7   # H3: time-driven tolerance


def test_incremental_algebra_equals_direct_restatement(golden):
    """deleting / appending points on the base solution gives the reference's set_waypoint + fit + predict
    numbers (golden cfg-3 sample, reference outputs) - the oracle of csrc/incremental.cu."""
    zl, z = golden('letters_ref.npz'), golden('cfg3_sample.npz')
    c = wl.CFG1
    t = wl.cfg1_kmp_grid(200)
    q = z['queries'][:, ::25]
    for r in (0, 3, 11, 17):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        a, b, cc = wl.cfg3_via_sets(li, zl['gmr_mu'][li], si + 1)
        xi, sg = orc.kmp_incremental(t, zl['gmr_mu'][li], zl['gmr_sigma'][li], list(a[si]), [b[si][j].reshape(1, -1) for j in range(2)],
                                     [cc[si][j] for j in range(2)], q, c['lam'], c['alpha'], c['sigma_f'], 0.0005)
        assert rel_err(xi, z['xi'][r][:, ::25]) < 1e-9 and rel_err(sg, z['sigma'][r][:, :, ::25]) < 1e-7, r
    # a via point that replaces a stored point with a slightly different input, and two via points on one stored point
    li = 3
    # ... and a second via point that only matches the stored point AFTER the first one moved it (in-place overwrite)
    for vs in ([t[0, 10] + 3e-4, 1.2345], [t[0, 50], t[0, 50] + 1e-4], [1.1495463763276121, 1.1490554230079142]):
        vp = [np.array([[1.0, -2.0]]), np.array([[0.5, 3.0]])]
        vv = [1e-4 * np.eye(2), np.array([[2e-3, 5e-4], [5e-4, 1e-3]])]
        s2, y2, g2 = orc.kmp_set_waypoint(t, zl['gmr_mu'][li], zl['gmr_sigma'][li], vs, vp, vv, 0.0005)
        est = orc.kmp_fit(s2, g2, c['lam'], c['sigma_f'], False)
        xr, sr = orc.kmp_predict(est, s2, y2, q, c['alpha'], c['sigma_f'], False)
        xi, sg = orc.kmp_incremental(t, zl['gmr_mu'][li], zl['gmr_sigma'][li], vs, vp, vv, q, c['lam'], c['alpha'], c['sigma_f'], 0.0005)
        assert rel_err(xi, xr) < 1e-9 and rel_err(sg, sr) < 1e-7
