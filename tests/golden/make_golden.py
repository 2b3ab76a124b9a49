"""Generates tests/golden/*.npz by running the UNMODIFIED reference (lbusellato/KMP_demos, mounted
read-only at /root/reference) in the authoring container.  The reference cannot travel to the GPU box,
so its outputs are committed as fixtures; this script is the recipe.

    python tests/golden/make_golden.py            # needs /root/reference, numpy, scipy, scikit-learn

Import shim: ``kmp/__init__.py`` imports matplotlib (absent here) and
``kmp/types/point.py`` fails on Python >= 3.11, so empty package modules and a stand-in ``Point`` are
registered before importing the reference's sub-modules.  Nothing of the reference is modified or copied.
"""
import os
import sys
import types
import warnings

import numpy as np

REF = os.environ.get('KMP_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def install_shim():
    kmp = types.ModuleType('kmp')
    kmp.__path__ = [REF + '/kmp']
    sys.modules['kmp'] = kmp
    t = types.ModuleType('kmp.types')
    t.__path__ = [REF + '/kmp/types']
    sys.modules['kmp.types'] = t
    p = types.ModuleType('kmp.types.point')

    class Point:
        pass
    Point.__module__ = 'kmp.types.point'
    p.Point = Point
    sys.modules['kmp.types.point'] = p


install_shim()
import logging  # noqa: E402
from kmp.model import KMP  # noqa: E402
from kmp.mixture import GaussianMixtureModel  # noqa: E402
from kmp.cluster import KMeans, Uniform  # noqa: E402
from kmp.types.quaternion import quaternion  # noqa: E402
from kmp import utils  # noqa: E402

logging.disable(logging.CRITICAL)
sys.path.insert(0, ROOT)
from kmp_demos_b200 import workloads as wl  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


def letters():
    """cfg 1 for all 26 letters: raw positions, reference GMM (K=8, reg 1e-6) and GMR."""
    pos_all, pri, mea, cov, nit, lb, gmu, gsig, labels0 = [], [], [], [], [], [], [], [], []
    for L in wl.LETTERS:
        demos = utils.read_struct(f'{REF}/2Dletters/{L}.mat', max_cell=5)
        pos = np.stack([d['pos'] for d in demos])            # (5,2,200)
        pos_all.append(pos)
        X = wl.cfg1_design_matrix(pos)
        g = GaussianMixtureModel(n_components=8, diag_reg_factor=1e-6)
        g.fit(X)
        mu, sig = g.predict(wl.cfg1_gmr_grid(200))
        pri.append(g.priors); mea.append(g.means); cov.append(g.covariances)
        nit.append(g.model.n_iter_); lb.append(g.model.lower_bound_)
        gmu.append(mu); gsig.append(sig)
        labels0.append(g.model.predict(X))
    save('letters_ref.npz', pos=np.stack(pos_all), priors=np.stack(pri), means=np.stack(mea),
         covariances=np.stack(cov), n_iter=np.array(nit), lower_bound=np.array(lb),
         gmr_mu=np.stack(gmu), gmr_sigma=np.stack(gsig), labels=np.stack(labels0))
    return np.stack(pos_all), np.stack(gmu), np.stack(gsig)


def cfg1(pos_all, gmu, gsig):
    """cfg 1 on G (full, N=200) and demo1's own ::4 variant (N=50): KMP fit/predict/2 via-points."""
    gi = wl.LETTERS.index('G')
    out = {}
    # --- N=200
    t = wl.cfg1_kmp_grid(200)
    k = KMP(l=0.5, alpha=40, sigma_f=200, time_driven_kernel=False)
    k.fit(t, gmu[gi], gsig[gi])
    q = wl.cfg1_query_grid(1000)
    xi0, sg0 = k.predict(q)
    vt = [0.5, 1.234]                                   # one on-grid (replace), one off-grid (append)
    vp = [np.array([[8.0, 4.0]]), np.array([[-3.0, -6.0]])]
    vs = [1e-4 * np.eye(2), 1e-4 * np.eye(2)]
    k.set_waypoint(vt, vp, vs)
    xi1, sg1 = k.predict(q)
    out.update(n200_pred_xi=xi0, n200_pred_sigma=sg0, n200_via_t=np.array(vt), n200_via_p=np.stack(vp),
               n200_via_var=np.stack(vs), n200_adapt_xi=xi1, n200_adapt_sigma=sg1, n200_adapt_s=k.s,
               n200_adapt_N=np.array(k.N))
    # --- demo1 proper: ::4 subsample
    pos = pos_all[gi]
    posc = np.concatenate([pos[h] for h in range(5)], axis=1)[:, ::4]
    N = posc.shape[1]
    T = N // 5
    time = 0.01 * np.tile(np.linspace(1, T, T), 5).reshape(1, -1)
    X = np.vstack((time, posc)).T
    g = GaussianMixtureModel(n_components=8, diag_reg_factor=1e-6)
    g.fit(X)
    mu, sig = g.predict(0.01 * np.arange(T).reshape(1, -1))
    tk = 0.01 * np.arange(1, T + 1).reshape(1, -1)
    k = KMP(l=0.5, alpha=40, sigma_f=200, time_driven_kernel=False)
    k.fit(tk, mu, sig)
    xi, sg = k.predict(tk)
    A = np.linalg.inv(k._estimator)
    out.update(d1_X=X, d1_priors=g.priors, d1_means=g.means, d1_covs=g.covariances,
               d1_n_iter=np.array(g.model.n_iter_), d1_lower_bound=np.array(g.model.lower_bound_),
               d1_gmr_mu=mu, d1_gmr_sigma=sig, d1_estimator=k._estimator, d1_matrix_from_inv=A,
               d1_pred_xi=xi, d1_pred_sigma=sg)
    k.set_waypoint([0.25], [np.array([[2.0, -1.0]])], [1e-4 * np.eye(2)])
    xi, sg = k.predict(tk)
    out.update(d1_adapt_xi=xi, d1_adapt_sigma=sg)
    save('cfg1_G.npz', **out)


def cfg3(gmu, gsig):
    """One reference-run cfg-3 problem per letter plus 8 extra on 'G' (M=100 queries to keep the file small)."""
    q = wl.cfg1_query_grid(1000)[:, ::10]
    t = wl.cfg1_kmp_grid(200)
    recs = []
    for li in range(26):
        n_sets = 9 if wl.LETTERS[li] == 'G' else 1
        vt, vp, vv = wl.cfg3_via_sets(li, gmu[li], n_sets)
        for s in range(n_sets):
            k = KMP(l=0.5, alpha=40, sigma_f=200, time_driven_kernel=False)
            k.fit(t, gmu[li], gsig[li])
            k.set_waypoint(list(vt[s]), [vp[s, j].reshape(1, -1) for j in range(2)], [vv[s, j] for j in range(2)])
            xi, sg = k.predict(q)
            recs.append((li, s, k.N, xi, sg))
    save('cfg3_sample.npz', letter=np.array([r[0] for r in recs]), set_idx=np.array([r[1] for r in recs]),
         N=np.array([r[2] for r in recs]), xi=np.stack([r[3] for r in recs]),
         sigma=np.stack([r[4] for r in recs]), queries=q)


def demo2():
    """cfg 2 (reference-faithful): two O=3 KMPs on pose_data.npy, incl. the un-listed-sigma quirk (H2)."""
    d = np.load(f'{REF}/quaternion_trajectories/pose_data.npy', allow_pickle=True)
    time = np.array([s.time for s in d]).reshape(1, -1)
    pos = np.vstack([s.pose[:3] for s in d]).T
    qe = np.vstack([s.quat_eucl for s in d]).T
    quat = np.vstack([s.quat.as_array() for s in d])
    demo_len = len(d) // 6
    ts = time[:, :demo_len]
    out = dict(time=time, pos=pos, quat_eucl=qe, quat=quat)
    for name, Y in (('pos', pos), ('quat', qe)):
        g = GaussianMixtureModel(n_demos=6)
        X = np.vstack((time, Y)).T
        g.fit(X)
        mu, sig = g.predict(ts)
        k = KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
        k.fit(ts, mu, sig)
        xi, sg = k.predict(ts)
        ka = KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
        ka.fit(ts, mu, sig)
        if name == 'pos':
            wp = np.array([0.0, 0.85, 0.3])
        else:
            qa = d[0].quat
            des = quaternion(-0.2, np.array([0.8, -0.7, -1.6]))
            wp = (des * (~qa)).log()
            out['des_quat'] = des.as_array()
            out['qa'] = qa.as_array()
        ka.set_waypoint([1], wp.reshape(1, -1), np.eye(3) * 1e-6)      # un-listed forms, demos.py:251,285
        xia, sga = ka.predict(ts)
        out.update({f'{name}_priors': g.priors, f'{name}_means': g.means, f'{name}_covs': g.covariances,
                    f'{name}_n_iter': np.array(g.model.n_iter_), f'{name}_gmr_mu': mu, f'{name}_gmr_sigma': sig,
                    f'{name}_pred_xi': xi, f'{name}_pred_sigma': sg, f'{name}_wp': wp,
                    f'{name}_adapt_xi': xia, f'{name}_adapt_sigma': sga, f'{name}_adapt_sigma_db': ka.sigma,
                    f'{name}_adapt_s': ka.s})
        if name == 'quat':
            kq = np.vstack((xia[:3, :], np.zeros_like(xia[0, :])))
            for i in range(kq.shape[1]):
                kq[:, i] = (quaternion.exp(kq[:3, i]) * qa).as_array()
            out['quat_adapt_back'] = kq
    # named variant: joint 7-D GMM, O=6
    g = GaussianMixtureModel(n_demos=6)
    X = np.vstack((time, pos, qe)).T
    g.fit(X)
    tq = np.linspace(0.01, 1.5, 400).reshape(1, -1)
    mu, sig = g.predict(tq)
    out.update(joint_priors=g.priors, joint_means=g.means, joint_covs=g.covariances,
               joint_n_iter=np.array(g.model.n_iter_), joint_gmr_mu=mu, joint_gmr_sigma=sig)
    tsub = tq[:, ::8]                                     # N=50, O=6 reference run (N=400 takes minutes)
    k = KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
    k.fit(tsub, mu[:, ::8], sig[:, :, ::8])
    xi, sg = k.predict(tq[:, ::16])
    out.update(joint_sub_pred_xi=xi, joint_sub_pred_sigma=sg)
    save('demo2_pose.npz', **out)


def demo3(pos_all):
    """demo3: time-driven kernel, O=4 (pos+vel), 4 via-points (demos.py:312-349)."""
    demos = utils.read_struct(f'{REF}/2Dletters/G.mat', max_cell=5)
    pos = np.concatenate([d['pos'] for d in demos], axis=1)
    vel = np.concatenate([d['vel'] for d in demos], axis=1)
    N = pos.shape[1]
    H = 5
    g = GaussianMixtureModel(n_demos=H)
    time = 0.01 * np.tile(np.linspace(1, N // H, N // H), H).reshape(1, -1)
    X = np.vstack((time, pos, vel)).T
    g.fit(X)
    mu, sig = g.predict(0.01 * np.arange(N // H).reshape(1, -1))
    tk = 0.01 * np.arange(1, N // H + 1).reshape(1, -1)
    k = KMP(l=1, sigma_f=6)
    k.fit(tk, mu, sig)
    xi0, sg0 = k.predict(tk[:, ::4])
    t = [0.01, 0.25, 1.2, 2]
    p = [np.array([8, 10, -50, 0.]).reshape(1, -1), np.array([-1, 6, -25, -40.]).reshape(1, -1),
         np.array([8, -4, 30, 10.]).reshape(1, -1), np.array([-3, 1, -10, 3.]).reshape(1, -1)]
    var = np.eye(4) * 1e-6
    k.set_waypoint(t, p, [var] * 4)
    xi1, sg1 = k.predict(tk)
    save('demo3_vel.npz', X=X, priors=g.priors, means=g.means, covs=g.covariances,
         n_iter=np.array(g.model.n_iter_), gmr_mu=mu, gmr_sigma=sig, pred_xi=xi0, pred_sigma=sg0,
         via_t=np.array(t), via_p=np.stack(p), adapt_xi=xi1, adapt_sigma=sg1, adapt_N=np.array(k.N))


def small_random():
    """Seeded small cases through the reference classes: kernel blocks, matrices, predict, GMR, KMeans,
    Uniform, quaternion maps."""
    rng = np.random.default_rng(12345)
    out = {}
    # kernel blocks + full matrices, I in {1,2}, plain and time-driven
    case = 0
    for (I, O, N, td, sf, lam) in [(1, 2, 7, False, 200.0, 0.5), (1, 4, 6, True, 6.0, 1.0), (2, 3, 5, False, 2.5, 0.3),
                                   (2, 2, 5, True, 1.5, 2.0), (1, 1, 9, False, 10.0, 0.1), (1, 6, 4, True, 3.0, 0.7)]:
        s = np.sort(rng.random((I, N)), axis=1) * 2
        mu = rng.normal(size=(O, N))
        Bm = rng.normal(size=(N, O, O))
        sig = np.transpose(np.einsum('nab,ncb->nac', Bm, Bm) + 0.1 * np.eye(O), (1, 2, 0)).copy()
        k = KMP(l=lam, alpha=13.0, sigma_f=sf, time_driven_kernel=td)
        k.fit(s, mu, sig)
        A = np.zeros((N * O, N * O))
        for i in range(N):
            for j in range(N):
                A[i * O:(i + 1) * O, j * O:(j + 1) * O] = k._KMP__kernel_matrix(s[:, i], s[:, j])
                if i == j:
                    A[i * O:(i + 1) * O, i * O:(i + 1) * O] += lam * sig[:, :, i]
        q = rng.random((I, 11)) * 2
        xi, sg = k.predict(q)
        out.update({f'kmp{case}_cfg': np.array([I, O, N, int(td), sf, lam, 13.0]), f'kmp{case}_s': s,
                    f'kmp{case}_mu': mu, f'kmp{case}_sigma': sig, f'kmp{case}_A': A,
                    f'kmp{case}_estimator': k._estimator, f'kmp{case}_q': q, f'kmp{case}_xi': xi,
                    f'kmp{case}_sg': sg, f'kmp{case}_kss': k._KMP__kernel_matrix(q[:, 0], q[:, 0])})
        case += 1
    out['kmp_cases'] = np.array(case)
    # GMR on random mixtures
    case = 0
    for (I, O, K, M, reg) in [(1, 2, 8, 40, 1e-6), (1, 4, 10, 30, 1e-4), (2, 3, 5, 25, 1e-4), (1, 6, 10, 20, 1e-4),
                              (1, 1, 3, 15, 1e-5)]:
        d = I + O
        g = GaussianMixtureModel(n_components=K, diag_reg_factor=reg)
        g.n_features = d
        pri = rng.random(K) + 0.1
        g.priors = pri / pri.sum()
        g.means = rng.normal(size=(d, K))
        Bm = rng.normal(size=(K, d, d)) * 0.5
        g.covariances = np.transpose(np.einsum('kab,kcb->kac', Bm, Bm) + 0.05 * np.eye(d), (1, 2, 0)).copy()
        x = rng.normal(size=(I, M))
        x[:, -1] = 1e3                                      # far-away query: all activations underflow
        mu, sig = g.predict(x)
        out.update({f'gmr{case}_cfg': np.array([I, O, K, M, reg]), f'gmr{case}_priors': g.priors,
                    f'gmr{case}_means': g.means, f'gmr{case}_covs': g.covariances, f'gmr{case}_x': x,
                    f'gmr{case}_mu': mu, f'gmr{case}_sigma': sig})
        case += 1
    out['gmr_cases'] = np.array(case)
    # reference KMeans (global RNG seeded) and Uniform
    case = 0
    for (d, n, K, seed) in [(3, 1000, 8, 1), (2, 257, 5, 2), (7, 600, 12, 3), (9, 300, 4, 4), (3, 64, 16, 5)]:
        X = rng.normal(size=(d, n)) + 3 * rng.integers(0, 3, size=(1, n))
        np.random.seed(seed)
        init_idx = np.random.choice(n, K, replace=False)
        np.random.seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            labels = KMeans(K).fit_predict(X)
        out.update({f'km{case}_X': X, f'km{case}_K': np.array(K), f'km{case}_seed': np.array(seed),
                    f'km{case}_init': init_idx, f'km{case}_labels': labels})
        case += 1
    out['km_cases'] = np.array(case)
    out['uniform_1000_8_5'] = Uniform(8, 5).fit_predict(np.zeros((3, 1000)))
    out['uniform_903_7_3'] = Uniform(7, 3).fit_predict(np.zeros((2, 903)))
    # quaternion maps
    Q = rng.normal(size=(64, 4))
    Q[0] = [1, 0, 0, 0]; Q[1] = [0, 1, 0, 0]; Q[2] = [-0.2, 0.8, -0.7, -1.6]; Q[3] = [-1, 1e-12, 0, 0]
    P = rng.normal(size=(64, 4))
    W = rng.normal(size=(64, 3))
    W[0] = 0; W[1] = 1e-9; W[2] = [2e-8, 0, 0]; W[3] = [0.508369, 0.857956, -0.167863]; W[4] = [3, 2, 1]
    qn, ql, qm, qc, qe, qr, qmc = [], [], [], [], [], [], []
    for i in range(64):
        q = quaternion(Q[i, 0], Q[i, 1:].copy())
        p = quaternion(P[i, 0], P[i, 1:].copy())
        qn.append(q.as_array()); ql.append(q.log()); qm.append((q * p).as_array()); qc.append((~q).as_array())
        qe.append(quaternion.exp(W[i]).as_array())
        qmc.append(((q * ~p)).log())
        with np.errstate(invalid='ignore', divide='ignore'):
            qr.append(quaternion.from_rotation_vector(W[i]).as_array())
    out.update(quat_Q=Q, quat_P=P, quat_W=W, quat_norm=np.stack(qn), quat_log=np.stack(ql), quat_mul=np.stack(qm),
               quat_conj=np.stack(qc), quat_exp=np.stack(qe), quat_rotvec=np.stack(qr), quat_mulconj_log=np.stack(qmc))
    save('small_cases.npz', **out)


def sklearn_gmm_synth():
    """sklearn GaussianMixture through the reference wrapper on cfg-5-shaped synthetic data (small n)."""
    X, _, _ = wl.cfg5_samples(20000, K=16, d=7, seed=3)
    g = GaussianMixtureModel(n_components=16, diag_reg_factor=1e-6)
    g.fit(X)
    save('gmm_synth.npz', seed=np.array(3), n=np.array(20000), K=np.array(16), priors=g.priors, means=g.means,
         covs=g.covariances, n_iter=np.array(g.model.n_iter_), lower_bound=np.array(g.model.lower_bound_),
         labels=g.model.predict(X), bic=np.array(g.bic(X)))


def sklearn_gmm_cfg5():
    """cfg 5 at its own K: the reference wrapper (sklearn GaussianMixture, random_state=420) on a 200 k-row draw of the
    cfg-5 generator, K=64, reg 1e-6 - the shape whose E-step sits on the 16-components-per-thread register boundary
    of the CUDA kernel.  Labels are stored as uint8 (K <= 256)."""
    import time
    X, _, _ = wl.cfg5_samples(200000, K=64, d=7, seed=0)
    g = GaussianMixtureModel(n_components=64, diag_reg_factor=1e-6)
    t0 = time.perf_counter()
    g.fit(X)
    sec = time.perf_counter() - t0
    save('gmm_cfg5_k64.npz', seed=np.array(0), n=np.array(200000), K=np.array(64), priors=g.priors, means=g.means,
         covs=g.covariances, n_iter=np.array(g.model.n_iter_), lower_bound=np.array(g.model.lower_bound_),
         converged=np.array(g.model.converged_), labels=g.model.predict(X).astype(np.uint8), bic=np.array(g.bic(X)),
         fit_seconds=np.array(sec), cores=np.array(len(os.sched_getaffinity(0))))


if __name__ == '__main__':
    parts = sys.argv[1:] or ['letters', 'cfg1', 'cfg3', 'demo2', 'demo3', 'small', 'gmm_synth', 'gmm_cfg5']
    if 'letters' in parts:
        pos_all, gmu, gsig = letters()
    else:
        z = np.load(os.path.join(HERE, 'letters_ref.npz'))
        pos_all, gmu, gsig = z['pos'], z['gmr_mu'], z['gmr_sigma']
    if 'cfg1' in parts:
        cfg1(pos_all, gmu, gsig)
    if 'cfg3' in parts:
        cfg3(gmu, gsig)
    if 'demo2' in parts:
        demo2()
    if 'demo3' in parts:
        demo3(pos_all)
    if 'small' in parts:
        small_random()
    if 'gmm_synth' in parts:
        sklearn_gmm_synth()
    if 'gmm_cfg5' in parts:
        sklearn_gmm_cfg5()
