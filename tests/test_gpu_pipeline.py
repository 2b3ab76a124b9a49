"""GPU parity of the reference-facing classes (kmp.model.KMP, kmp.mixture.GaussianMixtureModel, KMPBatch) against
golden outputs of the reference itself (tests/golden/*.npz) and the CPU oracle.

Tolerances (BASELINE.json north star): means <= 1e-9, covariances <= 1e-7 (norm-wise relative);
time-driven kernel <= 1e-7 on both."""
import os

import numpy as np
import pytest

from conftest import rel_err
from kmp_demos_b200 import workloads as wl

pytestmark = pytest.mark.gpu
MEAN_TOL, COV_TOL = 1e-9, 1e-7


def test_kmp_small_cases_vs_reference(golden):
    from kmp.model import KMP
    z = golden('small_cases.npz')
    for c in range(int(z['kmp_cases'])):
        I, O, N, td, sf, lam, alpha = z[f'kmp{c}_cfg']
        td = bool(td)
        k = KMP(l=lam, alpha=alpha, sigma_f=sf, time_driven_kernel=td)
        k.fit(z[f'kmp{c}_s'], z[f'kmp{c}_mu'], z[f'kmp{c}_sigma'])
        assert (k.O, k.N) == (int(O), int(N))
        xi, sg = k.predict(z[f'kmp{c}_q'])
        mt, ct = (1e-7, 1e-7) if td else (MEAN_TOL, COV_TOL)
        assert rel_err(xi, z[f'kmp{c}_xi']) < mt and rel_err(sg, z[f'kmp{c}_sg']) < ct, c
        assert rel_err(k.predict_mean(z[f'kmp{c}_q']), z[f'kmp{c}_xi']) < mt
        est = k._estimator
        assert rel_err(est, z[f'kmp{c}_estimator']) < (1e-6 if td else 1e-9), c
        assert rel_err(k.system_matrix(), z[f'kmp{c}_A']) < (1e-10 if td else 1e-15)


def test_cfg1_full_pipeline(golden):
    """cfg 1: G.mat -> GMM(8) -> GMR -> KMP(N=200) -> 2 via-points -> predict 1000."""
    from kmp.model import KMP
    from kmp.mixture import GaussianMixtureModel
    zl, z = golden('letters_ref.npz'), golden('cfg1_G.npz')
    gi = wl.LETTERS.index('G')
    X = wl.cfg1_design_matrix(zl['pos'][gi])
    gmm = GaussianMixtureModel(n_components=8, diag_reg_factor=1e-6)
    gmm.fit(X)
    assert gmm.model.n_iter_ == int(zl['n_iter'][gi])
    assert abs(gmm.model.lower_bound_ - float(zl['lower_bound'][gi])) < 1e-10
    assert rel_err(gmm.priors, zl['priors'][gi]) < MEAN_TOL and rel_err(gmm.means, zl['means'][gi]) < MEAN_TOL
    assert rel_err(gmm.covariances, zl['covariances'][gi]) < COV_TOL
    assert np.array_equal(gmm.model.predict(X), zl['labels'][gi])                 # argmax assignments: bit-exact
    mu, sig = gmm.predict(wl.cfg1_gmr_grid(200))
    assert rel_err(mu, zl['gmr_mu'][gi]) < MEAN_TOL and rel_err(sig, zl['gmr_sigma'][gi]) < COV_TOL
    c = wl.CFG1
    kmp = KMP(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
    kmp.fit(wl.cfg1_kmp_grid(200), mu, sig)
    q = wl.cfg1_query_grid(1000)
    xi, sg = kmp.predict(q)
    assert rel_err(xi, z['n200_pred_xi']) < MEAN_TOL and rel_err(sg, z['n200_pred_sigma']) < COV_TOL
    kmp.set_waypoint(list(z['n200_via_t']), [p for p in z['n200_via_p']], [v for v in z['n200_via_var']])
    assert kmp.N == int(z['n200_adapt_N']) and np.array_equal(kmp.s, z['n200_adapt_s'])
    xi, sg = kmp.predict(q)
    assert rel_err(xi, z['n200_adapt_xi']) < MEAN_TOL and rel_err(sg, z['n200_adapt_sigma']) < COV_TOL


def test_demo1_subsampled_variant(golden):
    from kmp.model import KMP
    from kmp.mixture import GaussianMixtureModel
    z = golden('cfg1_G.npz')
    gmm = GaussianMixtureModel(n_components=8, diag_reg_factor=1e-6)
    gmm.fit(z['d1_X'])
    assert gmm.model.n_iter_ == int(z['d1_n_iter']) == 33
    assert rel_err(gmm.means, z['d1_means']) < MEAN_TOL and rel_err(gmm.covariances, z['d1_covs']) < COV_TOL
    T = z['d1_gmr_mu'].shape[1]
    mu, sig = gmm.predict(0.01 * np.arange(T).reshape(1, -1))
    assert rel_err(mu, z['d1_gmr_mu']) < MEAN_TOL and rel_err(sig, z['d1_gmr_sigma']) < COV_TOL
    tk = 0.01 * np.arange(1, T + 1).reshape(1, -1)
    k = KMP(l=0.5, alpha=40, sigma_f=200, time_driven_kernel=False)
    k.fit(tk, z['d1_gmr_mu'], z['d1_gmr_sigma'])
    xi, sg = k.predict(tk)
    assert rel_err(xi, z['d1_pred_xi']) < MEAN_TOL and rel_err(sg, z['d1_pred_sigma']) < COV_TOL
    assert rel_err(k._estimator, z['d1_estimator']) < 1e-9
    k.set_waypoint([0.25], [np.array([[2.0, -1.0]])], [1e-4 * np.eye(2)])
    xi, sg = k.predict(tk)
    assert rel_err(xi, z['d1_adapt_xi']) < MEAN_TOL and rel_err(sg, z['d1_adapt_sigma']) < COV_TOL


@pytest.mark.parametrize('letter', ['A', 'M', 'S', 'Z'])
def test_gmm_fit_other_letters(golden, letter):
    from kmp.mixture import GaussianMixtureModel
    zl = golden('letters_ref.npz')
    li = wl.LETTERS.index(letter)
    X = wl.cfg1_design_matrix(zl['pos'][li])
    gmm = GaussianMixtureModel(n_components=8, diag_reg_factor=1e-6)
    gmm.fit(X)
    assert gmm.model.n_iter_ == int(zl['n_iter'][li])
    assert rel_err(gmm.means, zl['means'][li]) < MEAN_TOL and rel_err(gmm.covariances, zl['covariances'][li]) < COV_TOL
    assert np.array_equal(gmm.model.predict(X), zl['labels'][li])


def test_gmm_fit_synthetic_7d_and_bic(golden):
    from kmp.mixture import GaussianMixtureModel
    z = golden('gmm_synth.npz')
    X, _, _ = wl.cfg5_samples(int(z['n']), K=int(z['K']), d=7, seed=int(z['seed']))
    gmm = GaussianMixtureModel(n_components=int(z['K']), diag_reg_factor=1e-6)
    gmm.fit(X)
    assert gmm.model.n_iter_ == int(z['n_iter'])
    assert rel_err(gmm.priors, z['priors']) < MEAN_TOL and rel_err(gmm.means, z['means']) < MEAN_TOL
    assert rel_err(gmm.covariances, z['covs']) < COV_TOL
    assert np.array_equal(gmm.model.predict(X), z['labels'])
    assert abs(gmm.bic(X) - float(z['bic'])) < 1e-9 * abs(float(z['bic']))


def test_gmm_fit_cfg5_shape_k64(golden):
    """cfg 5 at its own K (64 components, d = 7 - the 16-components-per-thread boundary of k_gmm_em) on a 200 k-row
    draw: same iteration count as sklearn through the reference wrapper, labels bit-exact, parameters 1e-9 / 1e-7."""
    from kmp.mixture import GaussianMixtureModel
    z = golden('gmm_cfg5_k64.npz')
    X, _, _ = wl.cfg5_samples(int(z['n']), K=int(z['K']), d=7, seed=int(z['seed']))
    gmm = GaussianMixtureModel(n_components=int(z['K']), diag_reg_factor=1e-6)
    gmm.fit(X)
    assert gmm.model.n_iter_ == int(z['n_iter']) and gmm.model.converged_ == bool(z['converged'])
    assert abs(gmm.model.lower_bound_ - float(z['lower_bound'])) < 1e-9 * abs(float(z['lower_bound']))
    assert rel_err(gmm.priors, z['priors']) < MEAN_TOL and rel_err(gmm.means, z['means']) < MEAN_TOL
    assert rel_err(gmm.covariances, z['covs']) < COV_TOL
    # element-wise on the means where the reference value is not tiny
    big = np.abs(z['means']) > 1e-3
    assert np.max(np.abs(gmm.means - z['means'])[big] / np.abs(z['means'])[big]) < MEAN_TOL
    assert np.array_equal(gmm.model.predict(X), z['labels'].astype(np.int64))
    assert abs(gmm.bic(X) - float(z['bic'])) < 1e-9 * abs(float(z['bic']))


def test_kmeans_init_device_sampling_equals_host_sampling(golden):
    """k-means++ on the device (double-double prefix sums + provable-decision flags, kmpx_kpp_sample) picks the rows
    numpy's sequential cumsum / searchsorted picks: same centre indices, Lloyd iteration count and labels as the path
    that repeats every draw on the host with numpy's own expressions."""
    import torch
    from kmp_demos_b200.mixture import kmeans_init_labels
    for (n, K, seed) in ((20000, 16, 3), (5003, 5, 11), (257, 64, 2)):
        X, _, _ = wl.cfg5_samples(n, K=max(K, 8), d=7, seed=seed)
        Xd = torch.as_tensor(X, device='cuda')
        la, ia, _, fa = kmeans_init_labels(Xd, K, np.random.RandomState(420))
        lb, ib, _, fb = kmeans_init_labels(Xd, K, np.random.RandomState(420), host_sampling=True)
        assert torch.equal(fa['center_ids'], fb['center_ids']), (n, K)
        assert ia == ib and torch.equal(la, lb)
        assert fb['kpp_host_draws'] == K - 1 and fa['kpp_host_draws'] <= 1


def test_fused_fit_kernel_through_the_pipeline(golden):
    """KMPBatch.fused_fit = True (k_factor_struct instead of build + potrf + trtri) reproduces the reference's cfg-3
    results and agrees with the default pipeline to rounding."""
    from kmp_demos_b200.batch import KMPBatch
    zl, z = golden('letters_ref.npz'), golden('cfg3_sample.npz')
    c = wl.CFG1
    t = wl.cfg1_kmp_grid(200)
    base_of, vs, vp, vv = [], [], [], []
    for r in range(len(z['letter'])):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        a, b, cc = wl.cfg3_via_sets(li, zl['gmr_mu'][li], si + 1)
        base_of.append(li); vs.append(a[si]); vp.append(b[si]); vv.append(cc[si])
    res = []
    for fused in (False, True):
        kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False, chunk=16)
        kb.fused_fit = fused
        kb.set_base(np.broadcast_to(t, (26, 1, 200)), zl['gmr_mu'], zl['gmr_sigma'])
        res.append(kb.fit_predict(np.array(base_of), np.array(vs), np.array(vp), np.array(vv), z['queries']))
    for r in range(len(z['letter'])):
        assert rel_err(res[1][0][r], z['xi'][r]) < MEAN_TOL and rel_err(res[1][1][r], z['sigma'][r]) < COV_TOL, r
    assert rel_err(res[1][0], res[0][0]) < 1e-10 and rel_err(res[1][1], res[0][1]) < 1e-9


def test_warp_per_problem_fit_through_the_pipeline(golden):
    """KMPBatch.factor_mode = 'warp' (kmpx_potrf_wpp / kmpx_trtri_wpp with the two-output structure flag instead of the
    CTA-per-problem kernels) reproduces the reference's cfg-3 results, agrees with the 'cta' pipeline to rounding, ignores
    NaN-poisoned storage, and is independent of the chunking (a problem takes the same kernel in every chunk)."""
    from kmp_demos_b200.batch import KMPBatch
    zl, z = golden('letters_ref.npz'), golden('cfg3_sample.npz')
    c = wl.CFG1
    t = wl.cfg1_kmp_grid(200)
    base_of, vs, vp, vv = [], [], [], []
    for r in range(len(z['letter'])):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        a, b, cc = wl.cfg3_via_sets(li, zl['gmr_mu'][li], si + 1)
        base_of.append(li); vs.append(a[si]); vp.append(b[si]); vv.append(cc[si])
    args = (np.array(base_of), np.array(vs), np.array(vp), np.array(vv), z['queries'])
    res = {}
    for mode, chunk in (('cta', 16), ('warp', 16), ('warp', 7), ('persistent', 16)):
        kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False, chunk=chunk)
        if mode == 'persistent':            # the opt-in persistent predict kernel (predict_pers.cu) on top of the 'cta' fit
            kb.predict_mode = 'persistent'
        else:
            kb.factor_mode = mode
        kb.set_base(np.broadcast_to(t, (26, 1, 200)), zl['gmr_mu'], zl['gmr_sigma'])
        res[(mode, chunk)] = kb.fit_predict(*args)
        if (mode, chunk) == ('warp', 7):
            for key, buf in kb._bufs.items():
                if key[0] in ('A', 'w', 'alpha'):
                    buf.fill_(float('nan'))
            res['poisoned'] = kb.fit_predict(*args)
    xi, sg = res[('warp', 16)]
    for r in range(len(z['letter'])):
        assert rel_err(xi[r], z['xi'][r]) < MEAN_TOL and rel_err(sg[r], z['sigma'][r]) < COV_TOL, r
    assert rel_err(xi, res[('cta', 16)][0]) < 1e-10 and rel_err(sg, res[('cta', 16)][1]) < 1e-9
    assert np.array_equal(xi, res[('warp', 7)][0]) and np.array_equal(sg, res[('warp', 7)][1])
    assert np.array_equal(xi, res['poisoned'][0]) and np.array_equal(sg, res['poisoned'][1])
    xp, sp = res[('persistent', 16)]
    assert rel_err(xp, res[('cta', 16)][0]) < 1e-12 and rel_err(sp, res[('cta', 16)][1]) < 1e-11
    for r in range(len(z['letter'])):
        assert rel_err(xp[r], z['xi'][r]) < MEAN_TOL and rel_err(sp[r], z['sigma'][r]) < COV_TOL, r


def test_predict_kappa_box_skipping_stays_inside_rounding(golden):
    """The fused predict kernel drops boxes of 8 stored points whose kernel coefficients are below 2^-68 for a whole tile of
    32 or 64 queries (RBF locality; both CTA shapes are run).  Against the same kernel with every term kept (kmpx_predict_skip_threshold(0)): in-range queries agree to
    ~1e-16 norm-wise, and on a query grid that reaches far outside the data (tiles with no live box at all) mean and
    covariance still agree - far queries give mean 0 and covariance alpha * I up to the dropped 1e-21 terms."""
    from kmp_demos_b200 import _lib
    from kmp_demos_b200.batch import KMPBatch
    lib = _lib.load()
    zl, z = golden('letters_ref.npz'), golden('cfg3_sample.npz')
    c = wl.CFG1
    t = wl.cfg1_kmp_grid(200)
    base_of, vs, vp, vv = [], [], [], []
    for r in range(12):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        a, b, cc = wl.cfg3_via_sets(li, zl['gmr_mu'][li], si + 1)
        base_of.append(li); vs.append(a[si]); vp.append(b[si]); vv.append(cc[si])
    kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False, chunk=16)
    kb.set_base(np.broadcast_to(t, (26, 1, 200)), zl['gmr_mu'], zl['gmr_sigma'])
    for queries in (wl.cfg1_query_grid(1000), np.linspace(-6.0, 8.0, 1500).reshape(1, -1)):
        args = (np.array(base_of), np.array(vs), np.array(vp), np.array(vv), queries)
        xi1, sg1 = kb.fit_predict(*args)                       # default: boxes skipped, 32-query CTAs
        try:
            assert lib.kmpx_predict_tile_queries(64) == 0
            xi2, sg2 = kb.fit_predict(*args)                   # boxes skipped, 64-query CTAs
            assert lib.kmpx_predict_skip_threshold(0.0) == 0
            xi0, sg0 = kb.fit_predict(*args)                   # every term kept, 64-query CTAs
            assert lib.kmpx_predict_tile_queries(32) == 0
            xi3, sg3 = kb.fit_predict(*args)                   # every term kept, 32-query CTAs
        finally:
            assert lib.kmpx_predict_skip_threshold(2.0 ** -68) == 0 and lib.kmpx_predict_tile_queries(0) == 0
        for xa, sa in ((xi1, sg1), (xi2, sg2), (xi3, sg3)):
            assert rel_err(xa, xi0) < 1e-14 and rel_err(sa, sg0) < 1e-14
        assert np.all(np.isfinite(xi1)) and np.all(np.isfinite(sg1))
    far = np.abs(queries[0] - 1.0) > 3.0                      # further than any kernel support from the data in [0.01, 2]
    assert np.max(np.abs(xi1[:, :, far])) < 1e-15
    assert np.max(np.abs(sg1[:, 0, 0, far] - c['alpha'])) < 1e-12 and np.max(np.abs(sg1[:, 0, 1, far])) < 1e-12
    # a NaN query propagates as in the reference (its tile keeps every box), and only into its own outputs
    qn = np.array(queries, copy=True)
    qn[0, 5] = np.nan; qn[0, 1400] = np.nan                   # one in range of the data's tiles, one in a far tile
    xin, sgn = kb.fit_predict(np.array(base_of), np.array(vs), np.array(vp), np.array(vv), qn)
    bad = np.zeros(qn.shape[1], dtype=bool); bad[[5, 1400]] = True
    assert np.all(np.isnan(xin[:, :, bad])) and np.all(np.isnan(sgn[:, :, :, bad]))
    assert np.all(np.isfinite(xin[:, :, ~bad])) and rel_err(xin[:, :, ~bad], xi1[:, :, ~bad]) < 1e-14


def test_cfg3_batch_vs_reference_and_vs_single(golden):
    """cfg 3 sample: batched adaptation == the reference's per-problem results; batch == one-by-one bit for bit."""
    from kmp_demos_b200.batch import KMPBatch
    from kmp.model import KMP
    zl, z = golden('letters_ref.npz'), golden('cfg3_sample.npz')
    c = wl.CFG1
    t = wl.cfg1_kmp_grid(200)
    L = 26
    kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False, chunk=16)
    kb.set_base(np.broadcast_to(t, (L, 1, 200)), zl['gmr_mu'], zl['gmr_sigma'])
    base_of, vs, vp, vv = [], [], [], []
    for r in range(len(z['letter'])):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        a, b, cc = wl.cfg3_via_sets(li, zl['gmr_mu'][li], si + 1)
        base_of.append(li); vs.append(a[si]); vp.append(b[si]); vv.append(cc[si])
    xi, sg = kb.fit_predict(np.array(base_of), np.array(vs), np.array(vp), np.array(vv), z['queries'])
    for r in range(len(z['letter'])):
        assert rel_err(xi[r], z['xi'][r]) < MEAN_TOL and rel_err(sg[r], z['sigma'][r]) < COV_TOL, r
    # same problems through the single-problem class: same kernels => identical bits
    for r in (0, 7, 20):
        li = base_of[r]
        k = KMP(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
        k.fit(t, zl['gmr_mu'][li], zl['gmr_sigma'][li])
        k.set_waypoint(list(vs[r]), [vp[r][j].reshape(1, -1) for j in range(2)], [vv[r][j] for j in range(2)])
        assert k.N == int(z['N'][r])
        x1, s1 = k.predict(z['queries'])
        assert np.array_equal(x1, xi[r]) and np.array_equal(s1, sg[r])
    # sharding the batch (what a multi-GPU run does) leaves every problem's result bit-identical
    from kmp_demos_b200.dist import shard_range
    lo, hi = shard_range(len(base_of), 1, 2)
    x2, s2 = kb.fit_predict(np.array(base_of)[lo:hi], np.array(vs)[lo:hi], np.array(vp)[lo:hi], np.array(vv)[lo:hi], z['queries'])
    assert np.array_equal(x2, xi[lo:hi]) and np.array_equal(s2, sg[lo:hi])


def test_cfg3_batch_ignores_poisoned_storage(golden):
    """The strict upper triangle, the rows/columns beyond a system and the padding columns of the matrix buffer are
    never part of a result: with the buffer pre-filled with NaN the batched path still reproduces the reference
    (TMA tiles carry no mask; the kernels mask fragments in registers / rewrite what they stream)."""
    import torch
    from kmp_demos_b200.batch import KMPBatch
    zl, z = golden('letters_ref.npz'), golden('cfg3_sample.npz')
    c = wl.CFG1
    kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False, chunk=8)
    kb.set_base(np.broadcast_to(wl.cfg1_kmp_grid(200), (26, 1, 200)), zl['gmr_mu'], zl['gmr_sigma'])
    base_of, vs, vp, vv = [], [], [], []
    for r in range(len(z['letter'])):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        a, b, cc = wl.cfg3_via_sets(li, zl['gmr_mu'][li], si + 1)
        base_of.append(li); vs.append(a[si]); vp.append(b[si]); vv.append(cc[si])
    args = (np.array(base_of), np.array(vs), np.array(vp), np.array(vv), z['queries'])
    xi0, sg0 = kb.fit_predict(*args)
    for key, buf in kb._bufs.items():
        if key[0] in ('A', 'w', 'alpha'):
            buf.fill_(float('nan'))
    xi1, sg1 = kb.fit_predict(*args)
    assert np.array_equal(xi0, xi1) and np.array_equal(sg0, sg1)
    for r in range(len(z['letter'])):
        assert rel_err(xi1[r], z['xi'][r]) < MEAN_TOL and rel_err(sg1[r], z['sigma'][r]) < COV_TOL, r
    # pinned-output form of the device call: same bits as the device tensors
    dev = torch.device('cuda')
    d = [torch.as_tensor(np.ascontiguousarray(x), device=dev) for x in
         (np.array(base_of, dtype=np.int32), np.array(vs)[:, :, None], np.array(vp), np.array(vv), np.asarray(z['queries']).T)]
    B, M = len(base_of), z['queries'].shape[1]
    h_xi = torch.empty((B, 2, M), dtype=torch.float64).pin_memory()
    h_sg = torch.empty((B, 2, 2, M), dtype=torch.float64).pin_memory()
    xi_d, sg_d, info, _ = kb.run_device(*d, True, host_xi=h_xi, host_sigma=h_sg)
    torch.cuda.synchronize()
    assert int((info != 0).sum()) == 0
    assert torch.equal(h_xi, xi_d.cpu()) and torch.equal(h_sg, sg_d.cpu())
    assert np.array_equal(h_xi.numpy(), xi0)


def test_incremental_adaptation_equals_direct_and_reference(golden):
    """the low-rank correction of the base solution gives the reference's results (golden cfg-3 sample)
    and agrees with the direct batched path far inside the stated tolerances - including a via point that replaces
    a stored point with a slightly different input, two via points that hit the same stored point, one appended and
    one replaced point, and single-via-point problems."""
    import torch
    from kmp_demos_b200.batch import KMPBatch
    zl, z = golden('letters_ref.npz'), golden('cfg3_sample.npz')
    c = wl.CFG1
    grid = wl.cfg1_kmp_grid(200)
    kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False, chunk=16)
    kb.set_base(np.broadcast_to(grid, (26, 1, 200)), zl['gmr_mu'], zl['gmr_sigma'])
    base_of, vs, vp, vv = [], [], [], []
    for r in range(len(z['letter'])):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        a, b, cc = wl.cfg3_via_sets(li, zl['gmr_mu'][li], si + 1)
        base_of.append(li); vs.append(a[si]); vp.append(b[si]); vv.append(cc[si])
    n_gold = len(base_of)
    # hand-made edge cases on letter 3
    g = grid[0]
    extra = [(g[50], g[50] + 1e-4),          # both hit stored point 50: the later one supersedes the earlier one
             (g[10] + 3e-4, 1.2345),         # replace with a slightly different input, plus an append
             (g[199], g[0]),                 # two replacements at the ends of the trajectory
             (0.33333, 1.77777),             # two appends
             (1.1495463763276121, 1.1490554230079142)]   # the 2nd only matches point 114 after the 1st moved it (found by the 4-GPU bench)
    rng = np.random.default_rng(5)
    for t1, t2 in extra:
        base_of.append(3); vs.append(np.array([t1, t2])); vp.append(rng.normal(0.0, 3.0, (2, 2)))
        vv.append(np.stack([1e-4 * np.eye(2), np.array([[2e-3, 5e-4], [5e-4, 1e-3]])]))
    dev = torch.device('cuda')
    d = [torch.as_tensor(np.ascontiguousarray(x), device=dev) for x in
         (np.array(base_of, dtype=np.int32), np.array(vs)[:, :, None], np.array(vp), np.array(vv), np.asarray(z['queries']).T)]
    xd, sd, info, n_pts = kb.run_device(*d)
    xd, sd = xd.cpu().numpy(), sd.cpu().numpy()
    assert int((info != 0).sum()) == 0
    assert n_pts.cpu().numpy()[n_gold:].tolist() == [200, 201, 200, 202, 200]
    xi, sg, info_b = kb.run_incremental(*d)
    xi, sg = xi.cpu().numpy(), sg.cpu().numpy()
    assert int((info_b != 0).sum()) == 0
    for r in range(len(base_of)):
        assert rel_err(xi[r], xd[r]) < MEAN_TOL and rel_err(sg[r], sd[r]) < COV_TOL, r
    for r in range(n_gold):
        assert rel_err(xi[r], z['xi'][r]) < MEAN_TOL and rel_err(sg[r], z['sigma'][r]) < COV_TOL, r
    # one via point per problem
    d1 = [d[0], d[1][:, :1].contiguous(), d[2][:, :1].contiguous(), d[3][:, :1].contiguous(), d[4]]
    x1, s1, _, _ = kb.run_device(*d1)
    x2, s2, _ = kb.run_incremental(*d1)
    assert rel_err(x2.cpu().numpy(), x1.cpu().numpy()) < MEAN_TOL and rel_err(s2.cpu().numpy(), s1.cpu().numpy()) < COV_TOL


def test_incremental_adaptation_single_output():
    """O = 1 (scalar trajectory), odd number of stored points (identity padding row in the component-major system):
    incremental == direct."""
    import torch
    from kmp_demos_b200.batch import KMPBatch
    rng = np.random.default_rng(9)
    N0, L, B = 75, 3, 24
    t = np.linspace(0.02, 1.5, N0)
    mu = np.stack([np.sin(3.0 * t + l)[None] for l in range(L)])                        # (L,1,N0)
    sig = (0.01 + 0.05 * rng.random((L, 1, 1, N0)))
    kb = KMPBatch(l=0.7, alpha=10.0, sigma_f=150.0, chunk=8)
    kb.set_base(np.broadcast_to(t[None, None, :], (L, 1, N0)), mu, sig)
    base_of = rng.integers(0, L, B).astype(np.int32)
    via_s = np.where(rng.random((B, 2)) < 0.5, t[rng.integers(0, N0, (B, 2))], rng.uniform(0.02, 1.5, (B, 2)))
    via_s[:, 1] += 1e-3 * (np.abs(via_s[:, 1] - via_s[:, 0]) < 1e-3)
    via_p = rng.normal(0.0, 1.0, (B, 2, 1))
    via_v = np.full((B, 2, 1, 1), 1e-4)
    q = np.linspace(0.0, 1.6, 333)[:, None]
    d = [torch.as_tensor(np.ascontiguousarray(x), device='cuda') for x in (base_of, via_s[:, :, None], via_p, via_v, q)]
    xd, sd, info, _ = kb.run_device(*d)
    xd, sd = xd.cpu().numpy(), sd.cpu().numpy()
    assert int((info != 0).sum()) == 0
    xi, sg, _ = kb.run_incremental(*d)
    for r in range(B):
        assert rel_err(xi[r].cpu().numpy(), xd[r]) < MEAN_TOL and rel_err(sg[r].cpu().numpy(), sd[r]) < COV_TOL, r


def test_demo2_pose_and_unlisted_sigma_quirk(golden):
    """cfg 2 (reference-faithful): O=3 KMPs on pose_data, un-listed covariance => non-symmetric system (H2)."""
    from kmp.model import KMP
    from kmp.mixture import GaussianMixtureModel
    from kmp.types.quaternion import quaternion
    import kmp_demos_b200.types.quaternion as qm
    z = golden('demo2_pose.npz')
    time = z['time']
    ts = time[:, :150]
    for name, Y in (('pos', z['pos']), ('quat', z['quat_eucl'])):
        X = np.vstack((time, Y)).T
        gmm = GaussianMixtureModel(n_demos=6)
        gmm.fit(X)
        assert gmm.model.n_iter_ == int(z[f'{name}_n_iter'])
        assert rel_err(gmm.means, z[f'{name}_means']) < MEAN_TOL and rel_err(gmm.covariances, z[f'{name}_covs']) < COV_TOL
        mu, sig = gmm.predict(ts)
        assert rel_err(mu, z[f'{name}_gmr_mu']) < MEAN_TOL and rel_err(sig, z[f'{name}_gmr_sigma']) < COV_TOL
        k = KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
        k.fit(ts, z[f'{name}_gmr_mu'], z[f'{name}_gmr_sigma'])
        xi, sg = k.predict(ts)
        assert rel_err(xi, z[f'{name}_pred_xi']) < MEAN_TOL and rel_err(sg, z[f'{name}_pred_sigma']) < COV_TOL
        if name == 'pos':
            wp = np.array([0.0, 0.85, 0.3])
        else:
            qa = quaternion.from_array(z['qa'])
            des = quaternion(-0.2, np.array([0.8, -0.7, -1.6]))
            wp = (des * (~qa)).log()
            assert np.allclose(wp, z['quat_wp'], atol=1e-13)
        k.set_waypoint([1], wp.reshape(1, -1), np.eye(3) * 1e-6)          # demos.py:251,285
        assert np.array_equal(k.sigma, z[f'{name}_adapt_sigma_db'])
        assert k._fact.kind == 1, 'non-symmetric blocks must take the general path'
        xi, sg = k.predict(ts)
        assert rel_err(xi, z[f'{name}_adapt_xi']) < MEAN_TOL and rel_err(sg, z[f'{name}_adapt_sigma']) < COV_TOL
    back = qm.mul_batch(qm.exp_batch(z['quat_adapt_xi'][:3].T), z['qa'].reshape(1, 4))
    assert rel_err(back.T, z['quat_adapt_back']) < 1e-9
    # named variant: joint 7-D mixture, O = 6
    X = np.vstack((time, z['pos'], z['quat_eucl'])).T
    gmm = GaussianMixtureModel(n_demos=6)
    gmm.fit(X)
    assert gmm.model.n_iter_ == int(z['joint_n_iter'])
    tq = np.linspace(0.01, 1.5, 400).reshape(1, -1)
    mu, sig = gmm.predict(tq)
    assert rel_err(mu, z['joint_gmr_mu']) < MEAN_TOL and rel_err(sig, z['joint_gmr_sigma']) < COV_TOL
    k = KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
    k.fit(tq[:, ::8], z['joint_gmr_mu'][:, ::8], z['joint_gmr_sigma'][:, :, ::8])
    xi, sg = k.predict(tq[:, ::16])
    assert rel_err(xi, z['joint_sub_pred_xi']) < MEAN_TOL and rel_err(sg, z['joint_sub_pred_sigma']) < COV_TOL


def test_cfg2_named_variant_large_system():
    """O=6, N=400 (n = 2400 > KMPX_SMALL_MAX): the blocked large-n path against the vectorised oracle."""
    from kmp.model import KMP
    from oracle import kmp_oracle as orc
    rng = np.random.default_rng(11)
    N, O = 400, 6
    t = np.linspace(0.01, 1.5, N).reshape(1, -1)
    mu = np.stack([np.sin((o + 1) * t[0]) for o in range(O)])
    Bm = rng.normal(size=(N, O, O)) * 0.1
    sig = np.transpose(np.einsum('nab,ncb->nac', Bm, Bm) + 1e-3 * np.eye(O), (1, 2, 0)).copy()
    k = KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
    k.fit(t, mu, sig)
    q = np.linspace(0.0, 1.6, 77).reshape(1, -1)
    xi, sg = k.predict(q)
    est = orc.kmp_fit(t, sig, 0.5, 4.0, False)
    xr, sr = orc.kmp_predict(est, t, mu, q, 60.0, 4.0, False)
    assert rel_err(xi, xr) < MEAN_TOL and rel_err(sg, sr) < COV_TOL


def test_demo3_time_driven(golden):
    from kmp.model import KMP
    from kmp.mixture import GaussianMixtureModel
    z = golden('demo3_vel.npz')
    gmm = GaussianMixtureModel(n_demos=5)
    gmm.fit(z['X'])
    assert gmm.model.n_iter_ == int(z['n_iter'])
    assert rel_err(gmm.means, z['means']) < MEAN_TOL and rel_err(gmm.covariances, z['covs']) < COV_TOL
    mu, sig = gmm.predict(0.01 * np.arange(200).reshape(1, -1))
    assert rel_err(mu, z['gmr_mu']) < MEAN_TOL and rel_err(sig, z['gmr_sigma']) < COV_TOL
    tk = 0.01 * np.arange(1, 201).reshape(1, -1)
    k = KMP(l=1, sigma_f=6)
    k.fit(tk, z['gmr_mu'], z['gmr_sigma'])
    xi, sg = k.predict(tk[:, ::4])
    assert rel_err(xi, z['pred_xi']) < 1e-7 and rel_err(sg, z['pred_sigma']) < 1e-7       # H3 tolerance
    k.set_waypoint(list(z['via_t']), [p for p in z['via_p']], [np.eye(4) * 1e-6] * 4)
    assert k.N == int(z['adapt_N'])
    xi, sg = k.predict(tk)
    assert rel_err(xi, z['adapt_xi']) < 1e-7 and rel_err(sg, z['adapt_sigma']) < 1e-7


def test_cfg4_shape_reduced_and_properties():
    """cfg 4 at N=1024 (n=2048, blocked path, on-the-fly predict) against the oracle, plus size-independent
    properties at N=4096: zero posterior covariance is never negative, prediction at stored inputs with tiny
    lambda*Sigma reproduces mu."""
    from kmp.model import KMP
    from oracle import kmp_oracle as orc
    c = wl.CFG1
    t, mu, sig = wl.cfg4_problem(N=1024, seed=0)
    k = KMP(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
    k.fit(t, mu, sig)
    q = np.linspace(0.0, 2.0, 513).reshape(1, -1)
    xi, sg = k.predict(q)
    est = orc.kmp_fit(t, sig, c['lam'], c['sigma_f'], False)
    xr, sr = orc.kmp_predict(est, t, mu, q, c['alpha'], c['sigma_f'], False)
    assert rel_err(xi, xr) < MEAN_TOL and rel_err(sg, sr) < COV_TOL
    assert rel_err(k.predict_mean(q), xr) < MEAN_TOL
    t, mu, sig = wl.cfg4_problem(N=4096, seed=1)
    k = KMP(l=1e-3, alpha=1.0, sigma_f=c['sigma_f'], time_driven_kernel=False)
    k.fit(t, mu, sig)
    xi, sg = k.predict(t[:, ::16])
    assert rel_err(xi, mu[:, ::16]) < 5e-3                      # interpolation property for small lambda
    d0, d1 = sg[0, 0], sg[1, 1]
    assert np.all(d0 > -1e-6) and np.all(d1 > -1e-6) and np.all(d0 < 1.0 + 1e-6)


def test_cfg4_full_size_against_independent_lu_solve():
    """cfg 4 at BASELINE.json's full size (N = 8192, O = 2, n = 16384) through the class API.  The numpy oracle would
    need minutes for inv(A) here, so the checker is an independent fp64 LU route on the device (torch.linalg.solve on
    an A assembled in torch from the definition, KMP.py:146-156) - the same algebra the reference runs
    (k inv(A) Y, alpha (k** - k inv(A) k^T), KMP.py:187-190) - plus the size-independent identity
    L^-1 A L^-T = I on the stored factor."""
    import torch
    from kmp.model import KMP
    c = wl.CFG1
    N, O = 8192, 2
    t, mu, sig = wl.cfg4_problem(N=N, seed=3)
    k = KMP(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
    k.fit(t, mu, sig)
    M = 96
    q = np.linspace(-0.05, 2.05, M).reshape(1, -1)
    xi, sg = k.predict(q)
    dev = torch.device('cuda')
    td = torch.as_tensor(t[0], device=dev)
    n = N * O
    A = torch.zeros((n, n), dtype=torch.float64, device=dev)             # component-major: row = a*N + i
    Kd = torch.exp(-c['sigma_f'] * (td[:, None] - td[None, :]) ** 2)
    sd = torch.as_tensor(sig, device=dev)
    for a in range(O):
        A[a * N:(a + 1) * N, a * N:(a + 1) * N] = Kd
        for b in range(O):
            A[a * N:(a + 1) * N, b * N:(b + 1) * N] += torch.diag(c['lam'] * sd[a, b])
    del Kd
    f = k._fact
    assert f.kind == 0 and f.F.shape[0] == n
    Li = torch.tril(f.F[:, :n])
    R = Li @ A @ Li.T
    R.diagonal().sub_(1.0)
    ident = float(R.abs().max())
    del R, Li
    qd = torch.as_tensor(q[0], device=dev)
    ks = torch.exp(-c['sigma_f'] * (qd[:, None] - td[None, :]) ** 2)     # (M, N)
    rhs = torch.zeros((n, 1 + O * M), dtype=torch.float64, device=dev)
    rhs[:, 0] = torch.as_tensor(mu, device=dev).reshape(-1)              # Y in the component-major order
    for a in range(O):
        rhs[a * N:(a + 1) * N, 1 + a * M:1 + (a + 1) * M] = ks.T         # columns of k*^T for output a
    sol = torch.linalg.solve(A, rhs)
    xr = np.empty((O, M)); sr = np.empty((O, O, M))
    for a in range(O):
        xr[a] = (ks @ sol[a * N:(a + 1) * N, 0]).cpu().numpy()
        for b in range(O):
            quad = (ks * sol[a * N:(a + 1) * N, 1 + b * M:1 + (b + 1) * M].T).sum(1)
            sr[a, b] = (c['alpha'] * ((1.0 if a == b else 0.0) - quad)).cpu().numpy()
    assert ident < 1e-9, ident
    assert rel_err(xi, xr) < MEAN_TOL, rel_err(xi, xr)
    assert rel_err(sg, sr) < COV_TOL, rel_err(sg, sr)


@pytest.mark.parametrize('O', [1, 2])
def test_large_system_gemm_route_equals_streaming_kernel(O):
    """kmpx_kmp_predict_gemm (kappa + GEMMs + column reductions) against the fused streaming kernel on one n > 512
    system: several query chunks forced by a small workspace, M not a multiple of the chunk, odd N (padding row)."""
    import torch
    from kmp.model import KMP
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(21)
    N = 601
    t = np.linspace(0.0, 2.0, N).reshape(1, -1)
    mu = np.stack([np.sin(2.0 * t[0] + a) for a in range(O)])
    B = rng.normal(size=(N, O, O))
    sig = np.transpose(np.einsum('nab,ncb->nac', B, B) * 0.05 + 0.05 * np.eye(O), (1, 2, 0))
    k = KMP(l=0.5, alpha=3.0, sigma_f=150.0, time_driven_kernel=False)
    k.fit(t, mu, sig)
    f = k._fact
    n = f.F.shape[0]
    assert n > _lib.SMALL_MAX and f.kind == 0
    M = 300
    q = torch.as_tensor(np.linspace(-0.1, 2.1, M).reshape(-1, 1), device='cuda')
    xs = torch.empty((O, M), dtype=torch.float64, device='cuda'); ss = torch.empty((O, O, M), dtype=torch.float64, device='cuda')
    _lib.call('kmpx_kmp_predict', _lib.ptr(f.F), f.lda, n * f.lda, 0, _lib.ptr(f.w), f.lda, _lib.ptr(f.s_dev), None, N, 1, O, 1,
              _lib.ptr(q), 0, M, 150.0, 3.0, 0, _lib.ptr(xs), _lib.ptr(ss), _lib.stream())
    nbytes = int(lib.kmpx_kmp_predict_gemm_workspace_bytes(N, O, 128))            # room for chunks of 128 queries only
    ws = torch.empty(nbytes // 8, dtype=torch.float64, device='cuda')
    xg = torch.empty_like(xs); sg = torch.empty_like(ss)
    _lib.call('kmpx_kmp_predict_gemm', _lib.ptr(f.F), f.lda, _lib.ptr(f.w), _lib.ptr(f.s_dev), N, 1, O, _lib.ptr(q), M, 150.0, 3.0,
              _lib.ptr(ws), nbytes, _lib.ptr(xg), _lib.ptr(sg), _lib.stream())
    assert rel_err(xg.cpu().numpy(), xs.cpu().numpy()) < 1e-12 and rel_err(sg.cpu().numpy(), ss.cpu().numpy()) < 1e-11
    x1, s1 = k.predict(np.linspace(-0.1, 2.1, M).reshape(1, -1))                  # the class takes the GEMM route for n > 512
    assert rel_err(x1, xs.cpu().numpy()) < 1e-12 and rel_err(s1, ss.cpu().numpy()) < 1e-11
    with pytest.raises(_lib.KmpxError):
        _lib.call('kmpx_kmp_predict_gemm', _lib.ptr(f.F), f.lda, _lib.ptr(f.w), _lib.ptr(f.s_dev), N, 1, O, _lib.ptr(q), M, 150.0, 3.0,
                  _lib.ptr(ws), 1024, _lib.ptr(xg), _lib.ptr(sg), _lib.stream())


def test_batched_mean_only_path_equals_fused_path(golden):
    """KMPBatch with want_cov=False (alpha = A^-1 mu, then the mean-only predict kernel) against the fused
    mean+covariance path and the golden cfg-3 means."""
    import torch
    from kmp_demos_b200.batch import KMPBatch
    zl, z = golden('letters_ref.npz'), golden('cfg3_sample.npz')
    c = wl.CFG1
    kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False, chunk=8)
    kb.set_base(np.broadcast_to(wl.cfg1_kmp_grid(200), (26, 1, 200)), zl['gmr_mu'], zl['gmr_sigma'])
    base_of, vs, vp, vv = [], [], [], []
    for r in range(len(z['letter'])):
        li, si = int(z['letter'][r]), int(z['set_idx'][r])
        a, b, cc = wl.cfg3_via_sets(li, zl['gmr_mu'][li], si + 1)
        base_of.append(li); vs.append(a[si]); vp.append(b[si]); vv.append(cc[si])
    d = [torch.as_tensor(np.ascontiguousarray(x), device='cuda') for x in
         (np.array(base_of, dtype=np.int32), np.array(vs)[:, :, None], np.array(vp), np.array(vv), np.asarray(z['queries']).T)]
    x_full, _, info, _ = kb.run_device(*d, True)
    x_full = x_full.cpu().numpy()
    x_mean, sg_none, _, _ = kb.run_device(*d, False)
    assert sg_none is None and int((info != 0).sum()) == 0
    x_mean = x_mean.cpu().numpy()
    for r in range(len(base_of)):
        assert rel_err(x_mean[r], x_full[r]) < MEAN_TOL and rel_err(x_mean[r], z['xi'][r]) < MEAN_TOL, r


def _demo_data_root(tmp_path, golden):
    """Rebuilds the reference's data files (absent on the GPU box) from the committed fixtures: 2Dletters/G.mat as a
    1 x 5 cell array of structs (pos from letters_ref.npz, vel from the demo-3 design matrix) and
    quaternion_trajectories/pose_data.npy as pickled Point records."""
    import scipy.io as sio
    from kmp.types.point import Point
    from kmp.types.quaternion import quaternion
    pos = golden('letters_ref.npz')['pos'][6]                     # letter G: (5, 2, 200)
    X3 = golden('demo3_vel.npz')['X']                             # (1000, 5) = [t, pos, vel]
    cells = np.empty((1, 5), dtype=object)
    for h in range(5):
        blk = X3[h * 200:(h + 1) * 200]
        assert np.array_equal(blk[:, 1:3].T, pos[h])
        cells[0, h] = {'pos': pos[h], 'vel': blk[:, 3:5].T.copy(), 'acc': np.zeros((2, 200))}
    os.makedirs(tmp_path / '2Dletters')
    sio.savemat(str(tmp_path / '2Dletters' / 'G.mat'), {'demos': cells})
    z = golden('demo2_pose.npz')
    pts = []
    for i in range(z['time'].shape[1]):
        pose = np.zeros(6); pose[:3] = z['pos'][:, i]
        pts.append(Point(time=float(z['time'][0, i]), pose=pose, quat=quaternion.from_array(z['quat'][i]), quat_eucl=z['quat_eucl'][:, i].copy()))
    os.makedirs(tmp_path / 'quaternion_trajectories')
    arr = np.empty(len(pts), dtype=object)
    arr[:] = pts
    np.save(str(tmp_path / 'quaternion_trajectories' / 'pose_data.npy'), arr, allow_pickle=True)
    return str(tmp_path)


def test_headless_demos_reproduce_the_reference_pipelines(tmp_path, golden):
    """`from kmp import demo1, demo2, demo3` (main.py:12) without a display - every stage of the three demo
    pipelines against the outputs of the reference's own classes on the same files."""
    from kmp import demo1, demo2, demo3
    root = _demo_data_root(tmp_path, golden)
    z = golden('cfg1_G.npz')
    d = demo1(root=root)                                           # letter G, C = 8, every fourth sample
    assert d.gmm.model.n_iter_ == int(z['d1_n_iter'])
    assert rel_err(d.mu, z['d1_gmr_mu']) < MEAN_TOL and rel_err(d.sigma, z['d1_gmr_sigma']) < COV_TOL
    assert rel_err(d.mu_kmp, z['d1_pred_xi']) < MEAN_TOL and rel_err(d.sigma_kmp, z['d1_pred_sigma']) < COV_TOL
    t = d.add_waypoint(2.0, -1.0)                                  # ctrl-click: time of the closest GMR point
    dist = np.linalg.norm(d.mu[:2].T - np.array([2.0, -1.0]), axis=1)
    assert t == 0.01 * (int(np.argmin(dist)) + 1) and d.kmp.N == 50 + (0 if np.any(np.abs(d.time - t) < 5e-4) else 1)
    assert d.waypoints.shape == (1, 2) and set(d.results()) == set(demo1._fields)
    d.update()                                                     # "Update" resets the adaptation
    d.kmp.set_waypoint([0.25], [np.array([[2.0, -1.0]])], [1e-4 * np.eye(2)])
    xi, sg = d.kmp.predict(d.time)
    assert rel_err(xi, z['d1_adapt_xi']) < MEAN_TOL and rel_err(sg, z['d1_adapt_sigma']) < COV_TOL

    z = golden('demo3_vel.npz')
    d3 = demo3(root=root)
    assert d3.gmm.model.n_iter_ == int(z['n_iter'])
    assert rel_err(d3.mu, z['gmr_mu']) < MEAN_TOL and rel_err(d3.sigma, z['gmr_sigma']) < COV_TOL
    assert d3.kmp.N == int(z['adapt_N'])
    assert rel_err(d3.mu_kmp, z['adapt_xi']) < 1e-7 and rel_err(d3.sigma_kmp, z['adapt_sigma']) < 1e-7    # H3 tolerance
    out = d3.save(str(tmp_path / 'demo3.npz'))
    assert np.array_equal(np.load(out)['mu_kmp'], d3.mu_kmp)

    z = golden('demo2_pose.npz')
    d2 = demo2(root=root)
    assert d2.gmm.model.n_iter_ == int(z['pos_n_iter']) and d2.gmm_quat.model.n_iter_ == int(z['quat_n_iter'])
    assert rel_err(d2.mu_pos, z['pos_gmr_mu']) < MEAN_TOL and rel_err(d2.sigma_pos, z['pos_gmr_sigma']) < COV_TOL
    assert rel_err(d2.mu_pos_kmp, z['pos_pred_xi']) < MEAN_TOL and rel_err(d2.sigma_pos_kmp, z['pos_pred_sigma']) < COV_TOL
    assert rel_err(d2.mu_pos_kmp_ad, z['pos_adapt_xi']) < MEAN_TOL and rel_err(d2.sigma_pos_kmp_ad, z['pos_adapt_sigma']) < COV_TOL
    assert rel_err(d2.mu_quat_kmp, z['quat_pred_xi']) < MEAN_TOL and rel_err(d2.sigma_quat_kmp, z['quat_pred_sigma']) < COV_TOL
    assert np.allclose(d2.waypoint_quat, z['quat_wp'], atol=1e-13)
    assert rel_err(d2.mu_quat_kmp_ad, z['quat_adapt_xi']) < MEAN_TOL and rel_err(d2.sigma_quat_kmp_ad, z['quat_adapt_sigma']) < COV_TOL
    assert rel_err(d2.kmp_quats_ad, z['quat_adapt_back']) < 1e-9
    assert np.allclose(np.linalg.norm(d2.kmp_quats, axis=0), 1.0, atol=1e-12)
