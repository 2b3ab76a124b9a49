"""TEST / BASELINE INFRASTRUCTURE - CPU timings of the reference path for bench.py (SURVEY.md 8d "CPU baseline next to
it"): per configuration (i) the LITERAL reference (the unmodified classes from baseline/_ref via
oracle/ref_loader.py, else the oracle port) and, where BASELINE.md asks for it, (ii) a vectorised numpy/LAPACK
restatement as the "fair CPU" context line.  Bounded samples, every scaling rule stated in the result.  Never imported
by the product."""
import os
import time
import warnings

import numpy as np

from kmp_demos_b200 import workloads as wl
from . import kmp_oracle as orc
from . import ref_loader


def _quiet():
    import logging
    logging.disable(logging.CRITICAL)


def _classes():
    """(dict of classes, kind): the unmodified reference when it is installed, else None."""
    if ref_loader.available():
        _quiet()
        return ref_loader.classes(), 'reference'
    return None, 'port'


def cfg3_one_problem(args):
    """One cfg-3 problem = set_waypoint (2 via points: replace-or-append, then ONE full refit, KMP.py:57-91) + predict at
    M queries (KMP.py:160-194).  With the reference classes the model starts in the state `fit` leaves behind (s, xi,
    sigma, N, O attributes) WITHOUT paying for the base fit, so that both arms do one build + inverse per problem."""
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    t, mu, sig, vt, vp, vv, q, use_ref = args
    c = wl.CFG1
    xi_l = [vp[j].reshape(1, -1) for j in range(len(vt))]
    sg_l = [vv[j] for j in range(len(vt))]
    if use_ref:
        _quiet()
        KMP = ref_loader.load_module('KMP').KMP
        k = KMP(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
        k.s, k.xi, k.sigma = t.copy(), mu.copy(), sig.copy()
        k.O, k.N = mu.shape
        k.set_waypoint(list(vt), xi_l, sg_l)
        xi, _ = k.predict(q)
    else:
        s2, y2, g2 = orc.kmp_set_waypoint(t, mu, sig, list(vt), xi_l, sg_l, 5e-4)
        est = orc.kmp_fit(s2, g2, c['lam'], c['sigma_f'], False, literal=True)
        xi, _ = orc.kmp_predict_literal(est, s2, y2, q, c['alpha'], c['sigma_f'], False)
    return float(xi[0, 0])


def cfg1_line():
    """cfg 1 on one host core: GMM(8) fit -> GMR 200 -> KMP fit -> 2 via points (refit) -> predict (100 of the 1000
    steps, scaled x10: every query is an independent loop iteration, KMP.py:178)."""
    z = np.load(os.path.join(wl.GOLDEN_DIR, 'letters_ref.npz'))
    X = wl.cfg1_design_matrix(z['pos'][wl.LETTERS.index('G')])
    c = wl.CFG1
    cls, kind = _classes()
    via = ([0.5, 1.5], [np.array([[8.0, 4.0]]), np.array([[-3.0, -6.0]])], [1e-4 * np.eye(2)] * 2)
    q = wl.cfg1_query_grid(1000)[:, ::10]
    t = [time.perf_counter()]
    if cls:
        g = cls['GaussianMixtureModel'](n_components=8, diag_reg_factor=1e-6); g.fit(X); t.append(time.perf_counter())
        mu, sig = g.predict(wl.cfg1_gmr_grid(200)); t.append(time.perf_counter())
        k = cls['KMP'](l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
        k.fit(wl.cfg1_kmp_grid(200), mu, sig); t.append(time.perf_counter())
        k.set_waypoint(*via); t.append(time.perf_counter())
        k.predict(q); t.append(time.perf_counter())
    else:
        r = orc.gmm_fit(X, 8, 1e-6); t.append(time.perf_counter())
        pri, mea, cov = r['weights'], r['means'].T, np.transpose(r['covariances'], (1, 2, 0))
        mu, sig = orc.gmr_literal(pri, mea, cov, wl.cfg1_gmr_grid(200), 1e-6); t.append(time.perf_counter())
        s = wl.cfg1_kmp_grid(200)
        est = orc.kmp_fit(s, sig, c['lam'], c['sigma_f'], False, literal=True); t.append(time.perf_counter())
        s2, y2, g2 = orc.kmp_set_waypoint(s, mu, sig, *via, 5e-4)
        est = orc.kmp_fit(s2, g2, c['lam'], c['sigma_f'], False, literal=True); t.append(time.perf_counter())
        orc.kmp_predict_literal(est, s2, y2, q, c['alpha'], c['sigma_f'], False); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    return {'kind': kind, 'cores': 1, 'gmm_fit_ms': float(d[0]), 'gmr_ms': float(d[1]), 'kmp_fit_ms': float(d[2]),
            'set_waypoint_refit_ms': float(d[3]), 'predict_ms': float(d[4] * 10.0),
            'total_ms': float(d[:4].sum() + d[4] * 10.0), 'sample': 'predict timed on 100 of the 1000 steps, x10'}


def cfg2_line():
    """cfg 2, reference-faithful (demos.py:232-294): two O=3 KMPs (position, tangent-space orientation) on the 900
    samples of pose_data.npy (tests/golden/demo2_pose.npz carries the arrays): GMM(10) + GMR(150) + KMP fit + one via
    point (refit) + predict(150), per model.  One core."""
    z = np.load(os.path.join(wl.GOLDEN_DIR, 'demo2_pose.npz'))
    cls, kind = _classes()
    if not cls:
        return {'kind': 'unavailable', 'why': 'needs the reference classes (baseline/_ref)'}
    time_, ts = z['time'], z['time'][:, :150]
    out = {'kind': kind, 'cores': 1}
    tot = 0.0
    for name, Y, wp in (('pos', z['pos'], np.array([0.0, 0.85, 0.3])), ('quat', z['quat_eucl'], z['quat_wp'])):
        X = np.vstack((time_, Y)).T
        t = [time.perf_counter()]
        g = cls['GaussianMixtureModel'](n_demos=6); g.fit(X); t.append(time.perf_counter())
        mu, sig = g.predict(ts); t.append(time.perf_counter())
        k = cls['KMP'](l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
        k.fit(ts, mu, sig); t.append(time.perf_counter())
        k.set_waypoint([1], wp.reshape(1, -1), np.eye(3) * 1e-6); t.append(time.perf_counter())
        k.predict(ts); t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        out[name] = dict(zip(['gmm_fit_ms', 'gmr_ms', 'kmp_fit_ms', 'set_waypoint_refit_ms', 'predict_ms'], map(float, d)))
        tot += float(d.sum())
    out['total_ms'] = tot
    return out


def cfg4_lines(cores):
    """cfg 4 (n = 16384).  (ii) vectorised numpy/LAPACK at FULL size: broadcast exp build, scipy cho_factor (the
    "fair CPU" figure for the Cholesky TFLOP/s headline), alpha by cho_solve, posterior covariance for 256 queries by
    a triangular solve (scaled to 100 k).  (i) literal reference at N = 256 (n = 512), with the scaling law."""
    import scipy.linalg as sla
    N, O = 8192, 2
    c = wl.CFG1
    t, mu, sig = wl.cfg4_problem(N=N, seed=0)
    out = {}
    t0 = time.perf_counter()
    K = np.exp(-c['sigma_f'] * (t[0][:, None] - t[0][None, :]) ** 2)
    n = N * O
    A = np.empty((n, n), order='F')                                      # LAPACK works in place on column-major storage
    for a in range(O):
        for b in range(O):
            blk = A[a * N:(a + 1) * N, b * N:(b + 1) * N]
            if a == b:
                blk[...] = K
            else:
                blk[...] = 0.0
            blk[np.arange(N), np.arange(N)] += c['lam'] * sig[a, b]
    t1 = time.perf_counter()
    L = sla.cholesky(A, lower=True, overwrite_a=True, check_finite=False)
    t2 = time.perf_counter()
    y = mu.reshape(-1)                                                     # component-major right-hand side
    alpha = sla.cho_solve((L, True), y, check_finite=False)
    t3 = time.perf_counter()
    Mq = 256
    q = np.linspace(0.0, 2.0, Mq)
    ks = np.exp(-c['sigma_f'] * (t[0][:, None] - q[None, :]) ** 2)       # (N, Mq)
    rhs = np.zeros((n, O * Mq), order='F')
    for a in range(O):
        rhs[a * N:(a + 1) * N, a * Mq:(a + 1) * Mq] = ks
    V = sla.solve_triangular(L, rhs, lower=True, check_finite=False)
    cov = np.empty((O, O, Mq))
    for a in range(O):
        for b in range(O):
            cov[a, b] = c['alpha'] * ((1.0 if a == b else 0.0) - np.einsum('im,im->m', V[:, a * Mq:(a + 1) * Mq], V[:, b * Mq:(b + 1) * Mq]))
    mean = np.stack([ks.T @ alpha[a * N:(a + 1) * N] for a in range(O)])
    t4 = time.perf_counter()
    out['vectorised'] = {'kind': 'port (vectorised numpy + LAPACK; context line, not the reference)', 'cores': cores, 'n': n,
                         'build_ms': (t1 - t0) * 1e3, 'cholesky_ms': (t2 - t1) * 1e3,
                         'cholesky_tflops': n ** 3 / 3 / (t2 - t1) * 1e-12, 'solve_alpha_ms': (t3 - t2) * 1e3,
                         'predict_cov_ms_per_100k_queries': (t4 - t3) * 1e3 * 1e5 / Mq,
                         'sample': f'full-size build + Cholesky; predict on {Mq} queries, linear in M',
                         'check': {'mean0': float(mean[0, 1]), 'cov00': float(cov[0, 0, 1])}}
    del A, L, V, rhs, K
    cls, kind = _classes()
    Ns = 256
    ts, mus, sigs = wl.cfg4_problem(N=Ns, seed=0)
    qs = np.linspace(0.0, 2.0, 32).reshape(1, -1)
    t0 = time.perf_counter()
    if cls:
        k = cls['KMP'](l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
        k.fit(ts, mus, sigs); t1 = time.perf_counter()
        k.predict(qs); t2 = time.perf_counter()
    else:
        est = orc.kmp_fit(ts, sigs, c['lam'], c['sigma_f'], False, literal=True); t1 = time.perf_counter()
        orc.kmp_predict_literal(est, ts, mus, qs, c['alpha'], c['sigma_f'], False); t2 = time.perf_counter()
    ni = 2048
    Ai = np.random.default_rng(0).normal(size=(ni, ni)) + ni * np.eye(ni)
    t3 = time.perf_counter(); np.linalg.inv(Ai); t4 = time.perf_counter()
    fit_s, pred_s, inv_s = t1 - t0, (t2 - t1) / 32, t4 - t3
    out['literal'] = {'kind': kind, 'cores': cores, 'fit_s_at_N256': fit_s, 'predict_s_per_query_at_N256': pred_s,
                      'inv_s_at_n2048': inv_s,
                      'extrapolated_to_N8192': {'build_s': fit_s * (N / Ns) ** 2, 'inverse_s': inv_s * (n / ni) ** 3,
                                                'predict_s_per_100k_queries': pred_s * (N / Ns) ** 2 * 1e5,
                                                'law': 'build ~ N^2 kernel calls, LU inverse ~ n^3, predict ~ N kernel calls '
                                                       '+ 2 O n^2 flops per query'}}
    return out


def cfg5_line(cores, n=100_000, K=64):
    """cfg 5: what GaussianMixtureModel.fit executes underneath (sklearn GaussianMixture, full covariances, reg 1e-6)
    on an n-row draw of the cfg-5 generator; seconds per EM iteration from the difference of a 1-iteration and a
    4-iteration fit with a cheap initialisation (means_init = true means, init_params='random': no k-means), scaled
    linearly in n to 10 M rows."""
    from sklearn.mixture import GaussianMixture
    X, means, _ = wl.cfg5_samples(n, K=K, d=7, seed=0)
    times = []
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for it in (1, 4):
            gm = GaussianMixture(K, reg_covar=1e-6, random_state=420, max_iter=it, tol=0.0, init_params='random', means_init=means)
            t0 = time.perf_counter(); gm.fit(X); times.append(time.perf_counter() - t0)
    per_iter = (times[1] - times[0]) / 3.0
    return {'kind': 'reference (sklearn.mixture.GaussianMixture, the code GaussianMixtureModel.fit runs)', 'cores': cores,
            'n_sample': n, 'K': K, 's_per_iter_sample': per_iter, 's_per_iter_10M': per_iter * (1e7 / n),
            'sample': f'{n} rows, (4-iteration fit - 1-iteration fit) / 3, scaled x{1e7 / n:.0f}'}


def all_lines(cores):
    out = {}
    for name, fn in (('cfg1', cfg1_line), ('cfg2', cfg2_line), ('cfg4', lambda: cfg4_lines(cores)), ('cfg5', lambda: cfg5_line(cores))):
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:           # a CPU context line must never take the GPU bench down
            out[name] = {'kind': 'failed', 'why': repr(e)[:200]}
        out[name]['wall_s'] = time.perf_counter() - t0
    return out
