"""Per-kernel throughput on the B200: DGEMM, large Cholesky / triangular inverse, kernel build,
fused EM pass, GMR, quaternion maps, cfg-4 predict.  CUDA-event timing, best of `reps` after a warm-up.
`bench.py` reports the cfg-3 headline and embeds these figures under "extras";

    python -m kmp_demos_b200.perf [--quick] [--only dgemm,cfg4,em,gmr,quat]

prints them as one JSON object."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from . import _lib, workloads as wl

f64 = torch.float64


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    return json.load(open(p)).get('hbm_gbs', 6650.0) if os.path.isfile(p) else 6650.0


def dgemm_cublas(n=8192):
    a = torch.randn(n, n, dtype=f64, device='cuda'); b = torch.randn(n, n, dtype=f64, device='cuda'); c = torch.empty_like(a)
    ms = timeit(lambda: torch.matmul(a, b, out=c), reps=3, warm=2)
    return 2.0 * n ** 3 / ms * 1e-9


def dgemm_ours(n=8192, ta=0, tb=1, tri=0):
    a = torch.randn(n, n, dtype=f64, device='cuda'); b = torch.randn(n, n, dtype=f64, device='cuda'); c = torch.zeros(n, n, dtype=f64, device='cuda')
    st = _lib.stream()
    ms = timeit(lambda: _lib.call('kmpx_dgemm', ta, tb, n, n, n, 1.0, _lib.ptr(a), n, 0, _lib.ptr(b), n, 0, 0.0, _lib.ptr(c), n, 0, 1, tri, st))
    fl = 2.0 * n ** 3 * (0.5 if tri == 1 else 1.0)
    return fl / ms * 1e-9, ms


def cfg4(N, dg):
    lib = _lib.load()
    O = 2
    t, mu, sig = wl.cfg4_problem(N=N, seed=0)
    n = int(lib.kmpx_sys_size(N, O, _lib.LAYOUT_COMPONENT)); lda = n
    s_d = torch.as_tensor(np.ascontiguousarray(t.T), device='cuda')
    g_d = torch.as_tensor(np.ascontiguousarray(np.transpose(sig, (2, 0, 1))), device='cuda')
    y_d = torch.as_tensor(np.ascontiguousarray(mu.T), device='cuda')
    A = torch.empty((n, lda), dtype=f64, device='cuda')
    nbytes = max(int(lib.kmpx_potrf_workspace_bytes(n, 1)), int(lib.kmpx_trtri_workspace_bytes(n, 1))); ws = torch.empty(max(nbytes, 8) // 8, dtype=f64, device='cuda')
    info = torch.zeros(1, dtype=torch.int32, device='cuda'); st = _lib.stream(); c = wl.CFG1
    build = lambda uplo: _lib.call('kmpx_kernel_build', _lib.ptr(s_d), _lib.ptr(g_d), None, _lib.ptr(A), 1, N, 1, O, lda, n * lda,  # noqa: E731
                                   c['sigma_f'], c['lam'], 0, _lib.LAYOUT_COMPONENT, uplo, st)
    out = {}
    ms = timeit(lambda: build(_lib.UPLO_FULL))
    out['kernel_build_full'] = {'ms': ms, 'gbs': 8.0 * n * n / ms * 1e-6, 'frac_hbm': 8.0 * n * n / ms * 1e-6 / peaks()}
    ms = timeit(lambda: build(_lib.UPLO_LOWER))
    out['kernel_build_lower'] = {'ms': ms, 'gbs': 4.0 * n * (n + 1) / ms * 1e-6, 'frac_hbm': 4.0 * n * (n + 1) / ms * 1e-6 / peaks()}
    best = 1e30
    for _ in range(3):
        build(_lib.UPLO_LOWER)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, st)
        e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    assert int(info.item()) == 0
    out['potrf'] = {'n': n, 'ms': best, 'tflops': n ** 3 / 3 / best * 1e-9, 'frac_dgemm': n ** 3 / 3 / best * 1e-9 / dg}
    # the vendor library on the same matrix in the same run (torch.linalg.cholesky_ex -> cuSOLVER potrf; a measurement
    # aid, never on the product path): the figure kmpx_potrf has to match
    cus = 1e30
    for _ in range(3):
        build(_lib.UPLO_LOWER)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Lc, _ = torch.linalg.cholesky_ex(A[:, :n], check_errors=False)
        e1.record(); e1.synchronize(); cus = min(cus, e0.elapsed_time(e1))
    del Lc
    out['potrf'].update(cusolver_ms=cus, cusolver_tflops=n ** 3 / 3 / cus * 1e-9, vs_cusolver=cus / best)
    build(_lib.UPLO_LOWER)
    _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(ws), nbytes, st)
    e1.record(); e1.synchronize(); ms = e0.elapsed_time(e1)
    out['trtri'] = {'ms': ms, 'tflops': n ** 3 / 3 / ms * 1e-9, 'frac_dgemm': n ** 3 / 3 / ms * 1e-9 / dg}
    w = torch.zeros(lda, dtype=f64, device='cuda'); al = torch.empty((N, O), dtype=f64, device='cuda')
    ms = timeit(lambda: _lib.call('kmpx_solve_alpha', _lib.ptr(A), lda, n * lda, None, N, O, 1, _lib.LAYOUT_COMPONENT, _lib.ptr(y_d),
                                  _lib.ptr(w), lda, _lib.ptr(al), st), reps=2, warm=1)
    out['solve_alpha'] = {'ms': ms, 'gbs': 8.0 * n * n / ms * 1e-6}
    M = 8192
    q = torch.as_tensor(np.linspace(0, 2, M).reshape(-1, 1), device='cuda')
    xi = torch.empty((O, M), dtype=f64, device='cuda'); sg = torch.empty((O, O, M), dtype=f64, device='cuda')
    ms = timeit(lambda: _lib.call('kmpx_kmp_predict', _lib.ptr(A), lda, n * lda, 0, _lib.ptr(w), lda, _lib.ptr(s_d), None, N, 1, O, 1,
                                  _lib.ptr(q), 0, M, c['sigma_f'], c['alpha'], 0, _lib.ptr(xi), _lib.ptr(sg), st), reps=1, warm=1)
    fl = float(O * O) * N * N * M
    out['predict_cov'] = {'M': M, 'ms': ms, 'tflops': fl / ms * 1e-9, 'frac_dgemm': fl / ms * 1e-9 / dg, 'ms_per_100k_queries': ms * 1e5 / M}
    out['predict_cov_stream'] = out.pop('predict_cov')
    xi2 = torch.empty_like(xi); sg2 = torch.empty_like(sg)
    gbytes = int(lib.kmpx_kmp_predict_gemm_workspace_bytes(N, O, M)); gws = torch.empty(gbytes // 8, dtype=f64, device='cuda')
    ms = timeit(lambda: _lib.call('kmpx_kmp_predict_gemm', _lib.ptr(A), lda, _lib.ptr(w), _lib.ptr(s_d), N, 1, O, _lib.ptr(q), M,
                                  c['sigma_f'], c['alpha'], _lib.ptr(gws), gbytes, _lib.ptr(xi2), _lib.ptr(sg2), st), reps=2, warm=1)
    dev_x = float((xi2 - xi).norm() / xi.norm()); dev_s = float((sg2 - sg).norm() / sg.norm())
    out['predict_cov'] = {'M': M, 'route': 'kappa + k_dgemm + column reductions (predict_gemm.cu)', 'ms': ms, 'tflops': fl / ms * 1e-9,
                          'frac_dgemm': fl / ms * 1e-9 / dg, 'ms_per_100k_queries': ms * 1e5 / M,
                          'rel_dev_from_stream_kernel': {'mean': dev_x, 'covariance': dev_s}}
    del gws
    ms = timeit(lambda: _lib.call('kmpx_kmp_predict_mean', _lib.ptr(al), _lib.ptr(s_d), None, N, 1, O, 1, _lib.ptr(q), 0, M, c['sigma_f'], 0,
                                  _lib.ptr(xi), st), reps=2, warm=1)
    out['predict_mean'] = {'M': M, 'ms': ms, 'gexp_per_s': N * M / ms * 1e-6}
    return out


def em(n, K=64, d=7, dgemm_tflops=None):
    """Single-GPU fused EM pass on cfg-5 data (rows with their own nearest-mean initial labels)."""
    return em_partitioned(0, 1, torch.device('cuda', torch.cuda.current_device()), dgemm_tflops, n=n, K=K, d=d)


def cfg5_on_device(n, K, d, seed, dev):
    """cfg-5 samples drawn on the device: the mixture is wl.cfg5_mixture (numpy, seeded), the component ids and normal
    deviates come from torch's CUDA generator.  Returns X (n,d) and, as the initial hard assignment, the label of the
    nearest true mean OF EACH ROW (int32)."""
    means, A, _ = wl.cfg5_mixture(K, d, seed)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + seed)
    md, Ad = torch.as_tensor(means, device=dev), torch.as_tensor(A, device=dev)
    X = torch.empty((n, d), dtype=f64, device=dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    step = 1 << 20
    for c0 in range(0, n, step):
        c1 = min(n, c0 + step)
        comp = torch.randint(0, K, (c1 - c0,), generator=g, device=dev)
        z = torch.randn((c1 - c0, d), generator=g, dtype=f64, device=dev)
        X[c0:c1] = md[comp] + torch.einsum('nij,nj->ni', Ad[comp], z)
        lab[c0:c1] = torch.cdist(X[c0:c1], md).argmin(dim=1).to(torch.int32)
    return X, lab


def em_partitioned(rank, world, dev, dgemm_tflops, n=10_000_000, K=64, d=7, iters=20, warm=3):
    """cfg 5: EM over n rows split contiguously over `world` ranks (dist.shard_range), one NCCL all-reduce of the packed
    sufficient statistics per iteration.  Rank 0 draws the data and broadcasts it, so every rank slices the same rows.
    Timed: `iters` passes (E+M kernel, all-reduce, finalize) after `warm` passes, CUDA events, max over ranks; the
    all-reduce alone from events around torch.distributed.all_reduce.  At world > 1 rank 0 then repeats the same
    init + warm + iters passes on ALL rows alone: the one-GPU time of the same run (efficiency) and the parameters
    the partitioned run must reproduce (max deviation)."""
    import torch.distributed as dist
    from .mixture import EMEngine
    from . import dist as kd
    if rank == 0:
        X, lab = cfg5_on_device(n, K, d, 0, dev)
        shift = X[:1_000_000].mean(dim=0)
    else:
        X = torch.empty((n, d), dtype=f64, device=dev)
        lab = torch.empty(n, dtype=torch.int32, device=dev)
        shift = torch.empty(d, dtype=f64, device=dev)
    if world > 1:
        for t in (X, lab, shift):
            dist.broadcast(t, src=0)
    lo, hi = kd.shard_range(n, rank, world)

    def run(Xp, labp, group):
        eng = EMEngine(Xp, K, 1e-6, shift=shift, group=group, n_total=n)
        eng.init_from_labels(labp)
        for _ in range(warm):
            eng.em_pass()
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
            eng.allreduce_events = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            eng.em_pass()
        e1.record()
        e1.synchronize()
        return eng, e0.elapsed_time(e1) / iters

    eng, ms = run(X[lo:hi], lab[lo:hi], dist.group.WORLD if world > 1 else None)
    ms = kd.max_over_ranks(ms, dev)
    flops = 180.0 * n * K                                   # ~180 n K per iteration
    out = {'n': n, 'K': K, 'd': d, 'iters': iters, 'ranks': world, 'rows_per_rank': hi - lo, 'ms_per_iter': ms,
           'tflops_8d': flops / ms * 1e-9, 'frac_fp64': flops / ms * 1e-9 / (dgemm_tflops * world) if dgemm_tflops else None,
           'frac_hbm': 8.0 * n * d / ms * 1e-6 / (peaks() * world),
           'loglik_per_row': float(eng.stats[-1].item()) / n, 'info_max': int(eng.info.abs().max().item())}
    # GaussianMixtureModel.fit end to end on the same rows (k-means++ / Lloyd initialisation + EM to sklearn's
    # convergence rule), device times of the two stages
    from .mixture import GaussianMixtureModel
    gm = GaussianMixtureModel(n_components=K, diag_reg_factor=1e-6)
    gm.logger.disabled = True
    tm = {}
    gm.fit_device(X[lo:hi], group=dist.group.WORLD if world > 1 else None, n_total=n, n_before=lo, timings=tm)
    tm['init_share'] = tm['init_ms'] / (tm['init_ms'] + tm['em_ms'])
    tm['lower_bound'] = float(gm.model.lower_bound_)
    out['fit_end_to_end'] = tm
    if world > 1:
        us = sorted(a.elapsed_time(b) * 1e3 for a, b in eng.allreduce_events)
        out['allreduce_us_median'] = us[len(us) // 2]
        out['allreduce_us_min'] = us[0]
        out['allreduce_bytes'] = eng.stats.numel() * 8
        if rank == 0:
            ref, ms1 = run(X, lab, None)
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max())  # noqa: E731
            out['one_gpu_ms_per_iter_same_run'] = ms1
            out['efficiency_vs_one_gpu'] = ms1 / (world * ms)
            out['max_rel_dev_from_one_rank'] = {'weights': rel(eng.weights, ref.weights), 'means': rel(eng.means, ref.means),
                                                'covariances': rel(eng.covs, ref.covs)}
        dist.barrier()
    return out


def cfg4_sharded_predict(rank, world, dev, dgemm_tflops, N=8192, M=100_000):
    """KMP.predict of the n = 16384 system at M queries.  Rank 0 fits; with world > 1 the factor is
    broadcast over NCCL (KMP.broadcast_fit) and every rank predicts a contiguous slice of the queries, the slices are
    all-gathered on the device (KMP.predict_sharded_device).  Device times by CUDA events, max over ranks."""
    import torch.distributed as dist
    from .model import KMP
    from . import dist as kd
    c = wl.CFG1
    t, mu, sig = wl.cfg4_problem(N=N, seed=0)
    k = KMP(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
    k._logger.disabled = True
    out = {'n': 2 * N, 'M': M, 'ranks': world}
    if rank == 0:
        k.fit(t, mu, sig)
    sq = torch.as_tensor(np.linspace(0.0, 2.0, M).reshape(-1, 1), device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    if world > 1:
        dist.barrier()
        e0, e1 = ev(), ev()
        e0.record(); k.broadcast_fit(src=0); e1.record(); e1.synchronize()
        out['broadcast_fit_ms'] = kd.max_over_ranks(e0.elapsed_time(e1), dev)
        out['broadcast_bytes'] = int(k._fact.F.numel() * 8)
        k.predict_sharded_device(sq[:4096].contiguous())                  # warm-up: kernels loaded, workspace allocated
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = ev(), ev()
        e0.record(); res = k.predict_sharded_device(sq); e1.record(); e1.synchronize()
        out['predict_ms'] = kd.max_over_ranks(e0.elapsed_time(e1), dev)
    if rank == 0:
        k.predict_device(sq[:4096].contiguous())
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record(); xi, sg = k.predict_device(sq); e1.record(); e1.synchronize()
        out['one_gpu_predict_ms'] = e0.elapsed_time(e1)
        fl = 4.0 * N * N * M
        out['one_gpu_tflops'] = fl / out['one_gpu_predict_ms'] * 1e-9
        if world > 1:
            out['bit_identical_to_one_gpu'] = bool(torch.equal(res[:2], xi) and torch.equal(res[2:], sg.reshape(4, -1)))
            out['speedup'] = out['one_gpu_predict_ms'] / out['predict_ms']
            out['tflops_aggregate'] = fl / out['predict_ms'] * 1e-9
        else:
            out['predict_ms'] = out['one_gpu_predict_ms']
    if world > 1:
        dist.barrier()
    return out


def cfg2():
    """cfg 2 through the drop-in classes: (a) reference-faithful demo2 (two O = 3 KMPs, N = 150, K = 10) and (b) the
    variant BASELINE.json names (joint 7-D GMM, GMR on 400 points, one O = 6 KMP: n = 2400, via point, predict 400).
    Wall-clock per stage, host arrays in and out, best of three after a warm-up pass."""
    import time
    from .mixture import GaussianMixtureModel
    from .model import KMP
    z = np.load(os.path.join(wl.GOLDEN_DIR, 'demo2_pose.npz'))
    tm, ts = z['time'], z['time'][:, :150]
    names = ['gmm_fit_ms', 'gmr_ms', 'kmp_fit_ms', 'set_waypoint_refit_ms', 'predict_ms']

    def faithful():
        tot = np.zeros(5)
        for Y, wp in ((z['pos'], np.array([0.0, 0.85, 0.3])), (z['quat_eucl'], z['quat_wp'])):
            X = np.vstack((tm, Y)).T
            t = [time.perf_counter()]
            g = GaussianMixtureModel(n_demos=6); g.fit(X); torch.cuda.synchronize(); t.append(time.perf_counter())
            mu, sig = g.predict(ts); t.append(time.perf_counter())
            k = KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
            k.fit(ts, mu, sig); torch.cuda.synchronize(); t.append(time.perf_counter())
            k.set_waypoint([1], wp.reshape(1, -1), np.eye(3) * 1e-6); torch.cuda.synchronize(); t.append(time.perf_counter())
            k.predict(ts); t.append(time.perf_counter())
            tot += np.diff(t) * 1e3
        return tot

    def named():
        X = np.vstack((tm, z['pos'], z['quat_eucl'])).T
        tq = np.linspace(0.01, 1.5, 400).reshape(1, -1)
        t = [time.perf_counter()]
        g = GaussianMixtureModel(n_demos=6); g.fit(X); torch.cuda.synchronize(); t.append(time.perf_counter())
        mu, sig = g.predict(tq); t.append(time.perf_counter())
        k = KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)
        k.fit(tq, mu, sig); torch.cuda.synchronize(); t.append(time.perf_counter())
        k.set_waypoint([1.0], [mu[:, 266:267].T + 0.05], [np.eye(6) * 1e-6]); torch.cuda.synchronize(); t.append(time.perf_counter())
        k.predict(tq); t.append(time.perf_counter())
        return np.diff(t) * 1e3

    out = {}
    for key, fn in (('reference_faithful_2xO3_N150', faithful), ('named_O6_N400', named)):
        best = None
        for rep in range(4):
            d = fn()
            if rep > 0 and (best is None or d.sum() < best.sum()):
                best = d
        out[key] = {k_: float(v) for k_, v in zip(names, best)}
        out[key]['total_ms'] = float(best.sum())
    return out


def gmr(M, K=8, O=2):
    """GMR on M queries (K = 8, O = 2, scalar input: cfg 1's shape).  Two workloads: (a) the case SURVEY 8(d) names - a time
    GRID (linspace over the demonstrations' time span) through the reference's own cfg-1 mixture of letter G
    (tests/golden/letters_ref.npz): neighbouring queries activate the same 2-3 components; (b) worst case - unordered
    random queries through a random mixture whose components all overlap (nothing to skip, divergent warps)."""
    by = 8.0 * M * (1 + O + O * O)
    mu = torch.empty((O, M), dtype=f64, device='cuda'); sg = torch.empty((O, O, M), dtype=f64, device='cuda')

    def run(pri, mea, cov, x, reg):
        ms = timeit(lambda: _lib.call('kmpx_gmr', _lib.ptr(pri), _lib.ptr(mea), _lib.ptr(cov), K, 1, O, _lib.ptr(x), M, reg, _lib.ptr(mu),
                                      _lib.ptr(sg), _lib.stream()))
        return {'ms': ms, 'gbs': by / ms * 1e-6, 'frac_hbm': by / ms * 1e-6 / peaks()}

    z = np.load(os.path.join(wl.GOLDEN_DIR, 'letters_ref.npz'))
    gi = wl.LETTERS.index('G')
    dv = lambda a: torch.as_tensor(np.ascontiguousarray(a), device='cuda')  # noqa: E731
    grid = torch.linspace(0.0, 2.0, M, dtype=f64, device='cuda').reshape(1, M)
    out = {'M': M, 'K': K, 'O': O, 'grid_cfg1': run(dv(z['priors'][gi]), dv(z['means'][gi]), dv(z['covariances'][gi]), grid, 1e-6)}
    rng = np.random.default_rng(0)
    d = 1 + O
    Bm = rng.normal(size=(K, d, d))
    cov = dv(np.transpose(np.einsum('kab,kcb->kac', Bm, Bm) + 0.1 * np.eye(d), (1, 2, 0)))
    out['random_overlapping'] = run(dv(np.full(K, 1.0 / K)), dv(rng.normal(size=(d, K))), cov, torch.randn(1, M, dtype=f64, device='cuda'), 1e-6)
    out.update(ms=out['grid_cfg1']['ms'], gbs=out['grid_cfg1']['gbs'], frac_hbm=out['grid_cfg1']['frac_hbm'])
    return out


def quat(M):
    w = torch.randn(M, 3, dtype=f64, device='cuda'); q = torch.empty(M, 4, dtype=f64, device='cuda'); w2 = torch.empty_like(w)
    ms_e = timeit(lambda: _lib.call('kmpx_quat_exp', _lib.ptr(w), _lib.ptr(q), M, _lib.stream()))
    ms_l = timeit(lambda: _lib.call('kmpx_quat_log', _lib.ptr(q), _lib.ptr(w2), M, _lib.stream()))
    return {'M': M, 'exp_gbs': 56.0 * M / ms_e * 1e-6, 'log_gbs': 56.0 * M / ms_l * 1e-6, 'exp_frac_hbm': 56.0 * M / ms_e * 1e-6 / peaks(),
            'log_frac_hbm': 56.0 * M / ms_l * 1e-6 / peaks()}


def cfg1():
    """cfg 1 of BASELINE.json through the drop-in classes (the latency configuration: one trajectory, one model):
    GMM(8) fit -> GMR on 200 points -> KMP fit -> two via points -> predict 1000 steps.  Wall-clock per stage,
    host arrays in and out, best of three after one warm-up pass."""
    import time
    from .mixture import GaussianMixtureModel
    from .model import KMP
    z = np.load(os.path.join(wl.GOLDEN_DIR, 'letters_ref.npz'))          # the demonstrations of letter G travel with the fixture
    gi = wl.LETTERS.index('G')
    X = wl.cfg1_design_matrix(z['pos'][gi])
    c = wl.CFG1
    q = wl.cfg1_query_grid(1000)
    via = ([0.5, 1.5], [np.array([[8.0, 4.0]]), np.array([[-3.0, -6.0]])], [1e-4 * np.eye(2)] * 2)
    best = None
    for rep in range(4):
        t = [time.perf_counter()]
        gmm = GaussianMixtureModel(n_components=8, diag_reg_factor=1e-6); gmm.fit(X); torch.cuda.synchronize(); t.append(time.perf_counter())
        mu, sig = gmm.predict(wl.cfg1_gmr_grid(200)); t.append(time.perf_counter())
        kmp = KMP(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False)
        kmp.fit(wl.cfg1_kmp_grid(200), mu, sig); torch.cuda.synchronize(); t.append(time.perf_counter())
        kmp.set_waypoint(*via); torch.cuda.synchronize(); t.append(time.perf_counter())
        kmp.predict(q); t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        if rep > 0 and (best is None or d.sum() < best.sum()):
            best = d
    names = ['gmm_fit_ms', 'gmr_ms', 'kmp_fit_ms', 'set_waypoint_refit_ms', 'predict_ms']
    out = {k: float(v) for k, v in zip(names, best)}
    out['total_ms'] = float(best.sum())
    out['em_iterations'] = int(gmm.model.n_iter_)
    return out


def run_all(quick=False, only=None, dgemm_tflops=None):
    """All per-kernel figures as a dict (`only`: iterable of section names)."""
    only = set(only) if only else None
    want = lambda k: only is None or k in only  # noqa: E731
    _lib.load()
    out = {'gpu': torch.cuda.get_device_name(0), 'hbm_peak_gbs': peaks()}
    dg = dgemm_tflops or dgemm_cublas(4096 if quick else 8192)
    out['dgemm_cublas_tflops'] = dg
    if want('dgemm'):
        for name, (ta, tb, tri) in {'NT': (0, 1, 0), 'NN': (0, 0, 0), 'NT_lower': (0, 1, 1)}.items():
            tf, ms = dgemm_ours(4096 if quick else 8192, ta, tb, tri)
            out[f'dgemm_ours_{name}'] = {'tflops': tf, 'ms': ms, 'frac_cublas': tf / dg}
    if want('cfg1'):
        out['cfg1'] = cfg1()
    if want('cfg2'):
        out['cfg2'] = cfg2()
    if want('cfg4'):
        out['cfg4'] = cfg4(2048 if quick else 8192, dg)
    if want('em'):
        out['em'] = em(1_000_000 if quick else 10_000_000, dgemm_tflops=dg)
    if want('gmr'):
        out['gmr'] = gmr(1 << (22 if quick else 26))
    if want('quat'):
        out['quat'] = quat(1 << (22 if quick else 26))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    print(json.dumps(run_all(a.quick, a.only.split(',') if a.only else None)))


if __name__ == '__main__':
    main()
