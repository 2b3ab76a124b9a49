#!/usr/bin/env python
"""Benchmark of the KMP learn-and-adapt hot path (BASELINE.json metric: KMP fit+predict problems/sec, batched).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): cfg 3 of BASELINE.json - all 26 2Dletters reference trajectories (the reference's own
GMM+GMR output, tests/golden/letters_ref.npz) x 1024 seeded random via-point sets = 26 624 independent problems per
GPU; one problem = set_waypoint (2 via-points) -> fit (n = 400..404 fp64 system) -> predict mean + covariance at 1000
queries.  One "step" = one pass over the 26 624 problems of a rank.  Weak scaling: every rank gets its own 26 624
problems (different via-point seeds), no data-path collective.

`value`  : problems/s with inputs resident in HBM.
`e2e`    : the same through KMPBatch with HOST (pinned) buffers: via-points copied in, means+covariances copied out,
           every step, inside the timed region.
`roofline`: the dominant kernel (k_kmp_predict, FP64 tensor pipe) timed per launch with CUDA events inside the
           timed region; peak = cuBLAS DGEMM 8192^3 measured in the same run (MEASURED_PEAKS.json has no FP64 figure).
`strong` : the SAME 26 624 problems (rank 0's set) split over the N ranks by dist.shard_range - strong scaling of the
           named batch, results asserted bit-identical to the unsharded run.
`em`     : cfg 5 (10 M x 7, K = 64) row-partitioned over the N ranks, one NCCL all-reduce per iteration, at every N.
`extras` : per-kernel figures of SURVEY.md 8(d), the 100 k-query predict of the n = 16384 system sharded over the N
           ranks (8 f3), and (N = 1) the CPU lines of BASELINE.md section 4.
`--impl reference`: the reference's own classes (unmodified kmp/model/KMP.py from baseline/_ref; the oracle port when
           that copy is absent) on all host cores, a bounded sample of the same problems per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from kmp_demos_b200 import workloads as wl  # noqa: E402

T0 = time.perf_counter()


def log(msg):
    """Progress on stderr (stdout carries exactly one JSON line)."""
    sys.stderr.write(f'[bench +{time.perf_counter() - T0:7.1f}s] {msg}\n')
    sys.stderr.flush()


METRIC = 'KMP fit+predict problems/sec (batched)'
UNIT = 'problems/s'
N_SETS = 1024
M_QUERIES = 1000
WORKLOAD = 'cfg3: 26 letters x 1024 via-point sets, N=200..202, O=2, M=1000 queries, mean+covariance'


def build_problem_set(mu, sig, n_sets, seed_offset):
    L = mu.shape[0]
    ts, ps, vs = [], [], []
    for li in range(L):
        t, p, v = wl.cfg3_via_sets(li, mu[li], n_sets, seed_offset=seed_offset)
        ts.append(t); ps.append(p); vs.append(v)
    base_of = np.repeat(np.arange(L, dtype=np.int32), n_sets)
    return base_of, np.concatenate(ts)[:, :, None], np.concatenate(ps), np.concatenate(vs)


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_one_problem(args):
    """One cfg-3 problem through the reference's own classes (oracle/cpu_lines.py: unmodified KMP.set_waypoint -> fit ->
    predict when baseline/_ref is present, the oracle's literal port otherwise)."""
    from oracle import cpu_lines
    return cpu_lines.cfg3_one_problem(args)


def _cpu_kind():
    from oracle import ref_loader
    return 'reference' if ref_loader.available() else 'port'


def cpu_sample_args(mu, sig, base_of, via_s, via_p, via_v, count):
    t = wl.cfg1_kmp_grid(mu.shape[2])
    q = wl.cfg1_query_grid(M_QUERIES)
    idx = np.linspace(0, len(base_of) - 1, count).astype(int)       # stratified over the letters
    use_ref = _cpu_kind() == 'reference'
    return [(t, mu[base_of[i]], sig[base_of[i]], via_s[i, :, 0], via_p[i], via_v[i], q, use_ref) for i in idx]


def time_cpu(pool, sample):
    t0 = time.perf_counter()
    pool.map(_cpu_one_problem, sample, chunksize=1)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    mu, sig, src = wl.load_letter_references()
    base_of, via_s, via_p, via_v = build_problem_set(mu, sig, 8, 0)
    per_step = cores
    sample = cpu_sample_args(mu, sig, base_of, via_s, via_p, via_v, per_step)
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    with mp.get_context('spawn').Pool(cores) as pool:
        for _ in range(args.warmup):
            time_cpu(pool, sample)
        t = [time_cpu(pool, sample) for _ in range(args.steps)]
    sec = float(np.mean(t))
    value = per_step / sec
    kind = _cpu_kind()
    how = ("the reference's own KMP class (baseline/_ref/kmp/model/KMP.py, unmodified: set_waypoint -> fit -> predict)"
           if kind == 'reference' else "oracle port of KMP.set_waypoint/fit/predict (literal loops)")
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': src,
        'config': {'workload': WORKLOAD,
                   'step': f'bounded sample of the workload: {per_step} of its problems (one per host core) through {how}'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': f'{per_step} cfg-3 problems per step x {args.steps} steps, {how}, one process per core'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    _RESULT.append(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    def __init__(self, index):
        self.path = tempfile.mktemp(suffix='.csv')
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        for ln in open(self.path).read().splitlines():
            f = [x.strip() for x in ln.split(',')]
            if len(f) >= 7:
                try:
                    rows.append((float(f[0]), float(f[1]), float(f[2]), f[3:7]))
                except ValueError:
                    pass
        if rows:
            busy = [r for r in rows if r[2] > 0.5 * max(x[2] for x in rows)] or rows
            out['sm_mhz'] = float(np.median([r[0] for r in busy]))
            out['sm_max_mhz'] = rows[0][1]
            out['power_w_max'] = max(r[2] for r in rows)
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            out['reasons'] = [n for i, n in enumerate(names) if any(r[3][i].lower().startswith('active') for r in rows)]
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


def measure_dgemm_peak(torch, n=8192, seconds=1.5):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
    reps = max(3, int(seconds * 1e3 / e0.elapsed_time(e1)))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); e1.synchronize()
    return 2.0 * n ** 3 * reps / e0.elapsed_time(e1) * 1e-9


def extras_single_gpu(torch, dgemm_tflops):
    """The other measured rows of SURVEY.md 8(d) on one GPU (kmp_demos_b200/perf.py): second headline of BASELINE.json
    (fp64 Cholesky TFLOP/s at N*O = 16384), cfg-4 kernel build / triangular inverse / 100k-query predict rate,
    fused EM pass at 10 M x 7, K = 64, GMR and quaternion maps on 2^26 elements."""
    from kmp_demos_b200 import perf
    return perf.run_all(quick=False, only=('dgemm', 'cfg1', 'cfg2', 'cfg4', 'gmr', 'quat'), dgemm_tflops=dgemm_tflops)


def run_ours(args):
    import torch
    from kmp_demos_b200 import _lib, dist as kd
    from kmp_demos_b200.batch import KMPBatch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    rank, world, local = kd.init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    _lib.load()
    # pinned result buffers on the NUMA node of this rank's GPU (N > 1: all ranks copying to one node's memory is what
    # capped the 8-GPU end-to-end figure of the incremental path in round 1)
    numa_node = kd.bind_to_gpu_numa_node(local, world) if world > 1 else (None, 'single rank')
    log(f'rank {rank}: NUMA binding {numa_node}')
    mu, sig, src = wl.load_letter_references()
    L, O, T = mu.shape
    log(f'rank {rank}/{world}: building {L} x {args.sets} via-point sets')
    base_of, via_s, via_p, via_v = build_problem_set(mu, sig, args.sets, seed_offset=rank)
    B = len(base_of)
    cpu_line = cpu_other = None
    if rank == 0 and world == 1:
        # CPU baseline first (before the CUDA context exists in this process): oracle port on all host cores
        import multiprocessing as mp
        cores = len(os.sched_getaffinity(0))
        sample = cpu_sample_args(mu, sig, base_of, via_s, via_p, via_v, 2 * cores)
        os.environ['OPENBLAS_NUM_THREADS'] = '1'
        log(f'cpu_baseline: {len(sample)} problems on {cores} processes')
        with mp.get_context('spawn').Pool(cores) as pool:
            pool.map(_cpu_one_problem, sample[:cores], chunksize=1)          # warm-up: imports, page-in
            sec = time_cpu(pool, sample)
        os.environ.pop('OPENBLAS_NUM_THREADS', None)
        kind = _cpu_kind()
        cpu_line = {'value': len(sample) / sec, 'unit': UNIT, 'cores': cores, 'kind': kind,
                    'sample': f'{len(sample)} cfg-3 problems, '
                              + ("the reference's own KMP class (baseline/_ref, unmodified) set_waypoint -> fit -> predict"
                                 if kind == 'reference' else "oracle port of KMP.set_waypoint/fit/predict (literal loop nest)")
                              + f', one process per core, {sec:.1f} s wall'}
        if not args.no_extras:
            from oracle import cpu_lines
            log('cpu lines of the other configurations (BASELINE.md section 4)')
            cpu_other = cpu_lines.all_lines(cores)
        log(f'cpu_baseline done: {cpu_line["value"]:.2f} problems/s')
    c = wl.CFG1
    kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], time_driven_kernel=False, chunk=args.chunk)
    kb.set_base(np.broadcast_to(wl.cfg1_kmp_grid(T), (L, 1, T)), mu, sig)
    q_host = np.ascontiguousarray(wl.cfg1_query_grid(M_QUERIES).T)
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    h_base, h_vs, h_vp, h_vv, h_q = pin(base_of), pin(via_s), pin(via_p), pin(via_v), pin(q_host)
    d_base, d_vs, d_vp, d_vv, d_q = (x.to(dev) for x in (h_base, h_vs, h_vp, h_vv, h_q))
    xi = torch.empty((B, O, M_QUERIES), dtype=torch.float64, device=dev)
    sg = torch.empty((B, O, O, M_QUERIES), dtype=torch.float64, device=dev)
    h_xi = torch.empty(xi.shape, dtype=torch.float64).pin_memory()
    h_sg = torch.empty(sg.shape, dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_device():
        return kb.run_device(d_base, d_vs, d_vp, d_vv, d_q, True, xi, sg)

    def step_e2e():
        for dst, srcb in ((d_base, h_base), (d_vs, h_vs), (d_vp, h_vp), (d_vv, h_vv), (d_q, h_q)):
            dst.copy_(srcb, non_blocking=True)
        # results leave chunk by chunk on KMPBatch's copy stream while the next chunk is computed
        kb.run_device(d_base, d_vs, d_vp, d_vv, d_q, True, xi, sg, host_xi=h_xi, host_sigma=h_sg)

    def timed(fn, collect_events):
        for i in range(args.warmup):
            t1 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            log(f'{fn.__name__} warm-up {i}: {(time.perf_counter() - t1) * 1e3:.1f} ms')
        barrier()
        sampler = ClockSampler(local) if collect_events else None
        kb.events = [] if collect_events else None
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        launches = _lib.launch_count() - n0
        clocks = sampler.stop() if sampler else None
        events, kb.events = kb.events, None
        return kd.max_over_ranks(ms, dev), launches, clocks, events

    log(f'{B} problems per GPU resident; timing')
    ms_dev, launches, clocks, events = timed(step_device, True)
    log(f'device-resident: {ms_dev:.1f} ms/step')
    _, _, info, n_pts = step_device()
    torch.cuda.synchronize()
    assert int((info != 0).sum().item()) == 0, 'a system was not positive definite'
    n_host = n_pts.cpu().numpy().astype(np.float64)
    ms_e2e, _, _, _ = timed(step_e2e, False)
    log(f'end-to-end: {ms_e2e:.1f} ms/step')
    # sanity on the copied-out results: finite, covariance diagonals non-negative up to rounding
    assert bool(torch.isfinite(h_xi[:64]).all()) and float(h_sg[:64, 0, 0].min()) > -1e-6

    # ---- SURVEY 8(f1), reported NEXT TO the direct method (the metric, `value` and `roofline` stay the direct path):
    #      the same problems through the incremental adaptation, device-resident and end to end, and its deviation
    #      from the direct results of this very run
    def step_inc_device():
        return kb.run_incremental(d_base, d_vs, d_vp, d_vv, d_q, True, xi_inc, sg_inc)

    def step_inc_e2e():
        for dst, srcb in ((d_base, h_base), (d_vs, h_vs), (d_vp, h_vp), (d_vv, h_vv), (d_q, h_q)):
            dst.copy_(srcb, non_blocking=True)
        kb.run_incremental(d_base, d_vs, d_vp, d_vv, d_q, True, xi_inc, sg_inc, host_xi=h_xi, host_sigma=h_sg)

    xi_inc, sg_inc = torch.empty_like(xi), torch.empty_like(sg)
    step_device()
    ms_inc, _, _, _ = timed(step_inc_device, False)
    den_x = torch.linalg.norm(xi.reshape(B, -1), dim=1)
    den_s = torch.linalg.norm(sg.reshape(B, -1), dim=1)
    inc_err_mean = float((torch.linalg.norm((xi_inc - xi).reshape(B, -1), dim=1) / den_x).max())
    inc_err_cov = float((torch.linalg.norm((sg_inc - sg).reshape(B, -1), dim=1) / den_s).max())
    ms_inc_e2e, _, _, _ = timed(step_inc_e2e, False)
    log(f'incremental adaptation: {ms_inc:.1f} ms/step device-resident, {ms_inc_e2e:.1f} ms/step end to end; '
        f'max deviation from the direct path: mean {inc_err_mean:.1e}, covariance {inc_err_cov:.1e}')
    inc_ok = inc_err_mean < 1e-9 and inc_err_cov < 1e-7      # reported, never fatal: the metric is the direct method

    # ---- strong scaling of the NAMED batch (north star: "all 26 letters x 1024 via-point sets ... sharded across
    #      1/2/4/8 GPUs"): rank 0's 26 624 problems split contiguously by dist.shard_range, no collective in the data
    #      path.  Every rank also runs the unsharded batch once and asserts that its shard's results are the same bits.
    strong = None
    if world > 1:
        b0, s0, p0, v0 = build_problem_set(mu, sig, args.sets, seed_offset=0) if rank != 0 else (base_of, via_s, via_p, via_v)
        f_base, f_vs, f_vp, f_vv = (torch.as_tensor(np.ascontiguousarray(x), device=dev) for x in (b0, s0, p0, v0))
        kb.run_device(f_base, f_vs, f_vp, f_vv, d_q, True, xi, sg)              # unsharded results (xi, sg)
        lo, hi = kd.shard_range(B, rank, world)
        Bl = hi - lo
        l_base, l_vs, l_vp, l_vv = (x[lo:hi].contiguous() for x in (f_base, f_vs, f_vp, f_vv))
        hl_base, hl_vs, hl_vp, hl_vv = (pin(x.cpu().numpy()) for x in (l_base, l_vs, l_vp, l_vv))
        xi_l, sg_l = xi_inc[:Bl], sg_inc[:Bl]                                    # scratch of the right shape
        kb.chunk = strong_chunk(Bl, args.chunk)

        def step_strong_device():
            return kb.run_device(l_base, l_vs, l_vp, l_vv, d_q, True, xi_l, sg_l)

        def step_strong_e2e():
            for dst, srcb in ((l_base, hl_base), (l_vs, hl_vs), (l_vp, hl_vp), (l_vv, hl_vv), (d_q, h_q)):
                dst.copy_(srcb, non_blocking=True)
            kb.run_device(l_base, l_vs, l_vp, l_vv, d_q, True, xi_l, sg_l, host_xi=h_xi[:Bl], host_sigma=h_sg[:Bl])

        ms_s, _, _, _ = timed(step_strong_device, False)
        same = bool(torch.equal(xi_l, xi[lo:hi]) and torch.equal(sg_l, sg[lo:hi]))
        ms_se, _, _, _ = timed(step_strong_e2e, False)
        same_all = kd.max_over_ranks(0.0 if same else 1.0, dev) == 0.0
        kb.chunk = args.chunk
        log(f'strong scaling: {ms_s:.1f} ms/step device-resident, {ms_se:.1f} ms/step end to end, bit-identical: {same_all}')
        assert same_all, 'a sharded result differs from the unsharded run'
        strong = {'what': 'the same 26 624 problems (rank 0 set) split over the ranks by shard_range; no collective',
                  'problems_total': B, 'problems_per_gpu': Bl, 'chunk': strong_chunk(Bl, args.chunk),
                  'value': B / (ms_s * 1e-3), 'unit': UNIT, 'ms_per_step': ms_s,
                  'e2e': {'value': B / (ms_se * 1e-3), 'unit': UNIT, 'ms_per_step': ms_se},
                  'one_gpu_ms_per_step_same_run': ms_dev, 'speedup_vs_one_gpu': ms_dev / ms_s,
                  'efficiency': ms_dev / ms_s / world, 'e2e_efficiency': ms_e2e / ms_se / world,
                  'bit_identical_to_unsharded': same_all}
    else:
        strong = {'what': 'one GPU: the sharded batch is the whole batch', 'problems_total': B, 'problems_per_gpu': B,
                  'value': B / (ms_dev * 1e-3), 'unit': UNIT, 'ms_per_step': ms_dev,
                  'e2e': {'value': B / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e}, 'efficiency': 1.0}
    del xi_inc, sg_inc

    value = world * B / (ms_dev * 1e-3)
    e2e_value = world * B / (ms_e2e * 1e-3)
    h2d = sum(x.numel() * x.element_size() for x in (h_base, h_vs, h_vp, h_vv, h_q))
    d2h = h_xi.numel() * 8 + h_sg.numel() * 8
    del xi, sg, h_xi, h_sg
    torch.cuda.empty_cache()

    dgemm_tf = measure_dgemm_peak(torch)
    log(f'cuBLAS DGEMM peak {dgemm_tf:.2f} TFLOP/s')
    # ---- cfg 5: EM on row-partitioned samples with the NCCL all-reduce, at every N (all ranks take part)
    from kmp_demos_b200 import perf
    em_block = f3_block = None
    if not args.no_extras:
        log('cfg 5: partitioned EM')
        em_block = perf.em_partitioned(rank, world, dev, dgemm_tf)
        log(f'em: {em_block["ms_per_iter"]:.3f} ms/iter on {world} rank(s)')
        log('cfg 4: query-sharded predict of the n = 16384 system')
        f3_block = perf.cfg4_sharded_predict(rank, world, dev, dgemm_tf)
        torch.cuda.empty_cache()

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': src + '; via-point sets seeded per (letter, rank, set)',
        'config': {'workload': WORKLOAD,
                   'problems_per_gpu': B, 'chunk': args.chunk, 'numa_node_of_rank0': numa_node,
                   'l2': 'inputs larger than L2: each step streams ~35 GB of system matrices per GPU (126 MB L2)'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
        'gpu_launches': launches,
        'clocks': clocks,
    }
    if rank == 0:
        # per-kernel device time inside the timed region
        per = {}
        for name, a, b in events:
            per.setdefault(name, []).append(a.elapsed_time(b))
        kern = {k: {'ms_per_step': sum(v) / args.steps, 'launches_per_step': len(v) // args.steps} for k, v in per.items()}
        tot = sum(v['ms_per_step'] for v in kern.values())
        for v in kern.values():
            v['share'] = v['ms_per_step'] / tot
        line['kernels'] = kern
        pred = per.get('kmpx_kmp_predict_persistent') or per['kmpx_kmp_predict']
        n_launch = len(pred)
        # algorithmic flops of the fused predict: triangular V = Linv k*^T (O^2 Ns^2 M) + Gram/mean reduction
        Ns = np.ceil(n_host / 2) * 2
        flops_step = float(np.sum(O * O * Ns * Ns * M_QUERIES + 2.0 * O * Ns * (O * (O + 1) / 2 + O) * M_QUERIES))
        avg_ms = sum(pred) / n_launch
        achieved = flops_step * args.steps / n_launch / avg_ms * 1e-9
        # DRAM bytes per launch of the same kernel: STATIC figure from the committed ncu --set full capture under
        # profiles/ (not measured in this run), scaled to the problems per launch of this run; None when absent
        traffic, traffic_src = None, None
        for cand in ('r02_ncu_traffic.json', 'r01_ncu_traffic.json'):
            tpath = os.path.join(ROOT, 'profiles', cand)
            if os.path.isfile(tpath):
                with open(tpath) as fh:
                    tj = json.load(fh)
                    tk = tj.get('k_kmp_predict_tab') or tj['k_kmp_predict_fast']
                traffic = (tk['dram_bytes_read'] + tk['dram_bytes_write']) * (B * args.steps / n_launch) / tk['problems_per_launch']
                traffic_src = f'static: profiles/{cand} (ncu --set full of this kernel), scaled to the problems per launch'
                break
        # the predict kernel skips kappa boxes that are below 2^-68 for a whole query tile (RBF locality):
        # `achieved` / `frac` count the flops the tensor pipe EXECUTED (one extra, untimed step counts the k-steps on the
        # device); the dense algorithmic count of SURVEY 8(d) and its rate are reported next to it
        import ctypes
        lib_ = _lib.load()
        st_ = (ctypes.c_longlong * 2)()
        lib_.kmpx_debug_predict_stats(st_, 1)
        lib_.kmpx_debug_set(32768)
        nb_ = min(B, 2048)          # a sample of the batch: the executed fraction depends on N and the query grid only
        kb.run_device(d_base[:nb_], d_vs[:nb_], d_vp[:nb_], d_vv[:nb_], d_q, True)
        torch.cuda.synchronize()
        lib_.kmpx_debug_set(0)
        lib_.kmpx_debug_predict_stats(st_, 1)
        executed = st_[0] / st_[1] if st_[1] > 0 else 1.0
        line['roofline'] = {'kernel': 'k_kmp_predict', 'bound': 'tensor', 'achieved': achieved * executed, 'peak': dgemm_tf, 'unit': 'TFLOP/s',
                            'frac': achieved * executed / dgemm_tf, 'traffic': traffic, 'traffic_source': traffic_src, 'avg_launch_ms': avg_ms,
                            'flops_per_launch': flops_step * args.steps / n_launch * executed,
                            'executed_fraction_of_dense': executed,
                            'dense_algorithmic': {'flops_per_launch': flops_step * args.steps / n_launch, 'achieved': achieved,
                                                  'frac': achieved / dgemm_tf,
                                                  'note': 'O^2 Ns^2 M + reductions per problem (SURVEY 8d) / launch time: what a kernel without '
                                                          'the skipping would have to sustain for the same time'},
                            'peak_source': 'cuBLAS DGEMM 8192^3 sustained, measured in this run (FP64 tensor pipe)'}
        line['strong'] = strong
        line['em'] = em_block
        if world == 1:
            line['cpu_baseline'] = cpu_line
        if not args.no_extras:
            log('extras: per-kernel figures (cfg 1, cfg 2, cfg 4, GMR, quaternion maps)')
            line['extras'] = extras_single_gpu(torch, dgemm_tf)
            line['extras']['cfg4']['predict_100k_sharded'] = f3_block
            # second headline of BASELINE.json, promoted to a top-level key
            po = line['extras']['cfg4']['potrf']
            line['cholesky_n16384'] = {'metric': 'fp64 Cholesky TFLOP/s at N*O=16384', 'value': po['tflops'], 'unit': 'TFLOP/s',
                                       'ms': po['ms'], 'frac_dgemm': po['frac_dgemm'], 'cusolver_ms_same_run': po.get('cusolver_ms'),
                                       'cusolver_tflops_same_run': po.get('cusolver_tflops')}
            if cpu_other is not None:
                line['extras']['cpu_lines'] = cpu_other
        line['incremental'] = {
            'what': 'SURVEY 8(f1): same cfg-3 problems by exact low-rank correction of the per-base solution '
                    '(KMPBatch.run_incremental) instead of one refit per via-point set; bases re-factorised inside every step',
            'value': world * B / (ms_inc * 1e-3), 'unit': UNIT, 'ms_per_step': ms_inc,
            'e2e': {'value': world * B / (ms_inc_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_inc_e2e,
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'max_rel_dev_from_direct': {'mean': inc_err_mean, 'covariance': inc_err_cov}, 'within_tolerance': bool(inc_ok)}
        _RESULT.append(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def strong_chunk(b_local, chunk):
    """Systems resident at once for a shard of b_local problems: the factor kernels run twelve problems per SM (1776
    warp slots), so a shard is cut into near-equal chunks whose sizes are multiples of 1776 where the shard allows it
    (3 328 problems at N = 8 -> one chunk)."""
    if b_local <= chunk:
        return b_local
    n_chunks = -(-b_local // chunk)
    per = -(-b_local // n_chunks)
    return min(chunk, -(-per // 1776) * 1776)


_RESULT = []      # the JSON line(s) of this process, printed by main() once stdout is restored


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--sets', type=int, default=N_SETS, help='via-point sets per letter (1024 = the named config)')
    ap.add_argument('--chunk', type=int, default=5328,
                    help='systems resident at once (a multiple of 12 x 148 = 1776: the factor kernels run twelve problems per SM)')
    ap.add_argument('--no-extras', action='store_true')
    args = ap.parse_args()
    # stdout carries exactly one JSON line: whatever libraries write to file descriptor 1 while the run is on (NCCL's
    # version banner under NCCL_DEBUG=VERSION, for one) is sent to stderr, and the descriptor is restored for the result
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    _RESULT.clear()
    try:
        if args.impl == 'reference':
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for text in _RESULT:
        print(text, flush=True)


if __name__ == '__main__':
    main()
