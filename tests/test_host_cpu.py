"""CPU-only tests: the C-ABI library loads and exports every symbol include/kmpx.h declares (no compute calls),
host-side logic (sharding, workloads, argument validation that happens before any launch), and the N>1 path on a
world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from kmp_demos_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        from kmp_demos_b200.csrc import build
        build.build()
    header = open(os.path.join(ROOT, 'include', 'kmpx.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(kmpx_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in include/kmpx.h but not exported'
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    assert _lib.load().kmpx_version() == 100


def test_argument_validation_without_gpu():
    """Bad arguments are rejected on the host before any launch: negative status + message."""
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    rc = lib.kmpx_kernel_build(None, None, None, None, 1, 4, 1, 2, 8, 64, 1.0, 0.5, 0, 0, 0, None)
    assert rc == -1 and b'NULL' in lib.kmpx_last_error()
    rc = lib.kmpx_dgemm(0, 0, 4, 4, 4, 1.0, ctypes.c_void_p(16), 3, 0, ctypes.c_void_p(16), 4, 0, 0.0, ctypes.c_void_p(16), 4, 0, 1, 0, None)
    assert rc == -8
    assert lib.kmpx_sys_size(201, 2, _lib.LAYOUT_COMPONENT) == 404 and lib.kmpx_sys_size(201, 2, _lib.LAYOUT_POINT) == 402
    assert lib.kmpx_potrf_workspace_bytes(512, 1) == 0 and lib.kmpx_potrf_workspace_bytes(16384, 1) > 0
    # process-wide options of the fused predict kernel: range-checked, and restored to their defaults here
    assert lib.kmpx_predict_tile_queries(48) == -1 and b'32 or 64' in lib.kmpx_last_error()
    assert lib.kmpx_predict_tile_queries(32) == 0 and lib.kmpx_predict_tile_queries(64) == 0 and lib.kmpx_predict_tile_queries(0) == 0
    assert lib.kmpx_predict_skip_threshold(1e-3) == -1 and lib.kmpx_predict_skip_threshold(float('nan')) == -1
    assert lib.kmpx_predict_skip_threshold(2.0 ** -68) == 0


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from kmp.model import KMP
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        KMP().fit(np.zeros((1, 3)), np.zeros((2, 3)), np.zeros((2, 2, 3)))
    from kmp.types.quaternion import quaternion
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        quaternion(1.0, np.zeros(3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'kmp_demos_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('no oracle import', ''), f'{f} mentions the oracle'


def test_shard_range_partitions_exactly():
    from kmp_demos_b200.dist import shard_range
    for n in (0, 1, 7, 26624, 26625):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_workloads_are_deterministic_and_shaped():
    from kmp_demos_b200 import workloads as wl
    mu, sig, src = wl.load_letter_references()
    assert mu.shape == (26, 2, 200) and sig.shape == (26, 2, 2, 200)
    t1, p1, v1 = wl.cfg3_via_sets(3, mu[3], 5)
    t2, p2, v2 = wl.cfg3_via_sets(3, mu[3], 9)
    assert np.array_equal(t1, t2[:5]) and np.array_equal(p1, p2[:5])
    assert np.all(t1 >= 0.01) and np.all(t1 <= 2.0) and v1.shape == (5, 2, 2, 2)
    t, m, s = wl.cfg4_problem(N=64)
    assert np.all(np.linalg.eigvalsh(np.transpose(s, (2, 0, 1))) > 0.04)
    X, _, _ = wl.cfg5_samples(1000, K=8, d=7)
    assert X.shape == (1000, 7) and X.flags.c_contiguous


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["KMP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from kmp_demos_b200 import dist as kd, workloads as wl
from kmp_demos_b200.mixture import stats_len
from oracle import kmp_oracle as orc
rank, world, _ = kd.init_process_group("gloo")
X, means, _ = wl.cfg5_samples(4000, K=6, d=4, seed=2)
lab = np.argmin(((X[:, None, :] - means[None]) ** 2).sum(-1), axis=1)
lo, hi = kd.shard_range(X.shape[0], rank, world)
# per-rank sufficient statistics of the one-hot M-step in the packed layout of kmpx_gmm_label_stats
K, d = 6, 4
F = 1 + d + d * (d + 1) // 2
st = np.zeros(stats_len(K, d))
iu = np.triu_indices(d)
for k in range(K):
    xs = X[lo:hi][lab[lo:hi] == k]
    st[k * F] = len(xs)
    st[k * F + 1:k * F + 1 + d] = xs.sum(0)
    st[k * F + 1 + d:(k + 1) * F] = (xs[:, :, None] * xs[:, None, :]).sum(0)[iu]
t = torch.from_numpy(st)
kd.allreduce_stats(t)
tot = t.numpy().reshape(-1)[:-1].reshape(K, F)
nk = tot[:, 0]
mu = tot[:, 1:1 + d] / nk[:, None]
resp = np.zeros((X.shape[0], K)); resp[np.arange(X.shape[0]), lab] = 1
nk_ref, mu_ref, _ = orc.gmm_m_step(X, resp, 1e-6)
assert np.allclose(nk, nk_ref - 10 * np.finfo(float).eps) and np.allclose(mu, mu_ref, rtol=1e-12, atol=1e-12)
ms = kd.max_over_ranks(10.0 + rank, torch.device("cpu"))
assert ms == 10.0 + world - 1
dist.barrier()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_allreduce_of_statistics(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, KMP_ROOT=ROOT, MASTER_ADDR='127.0.0.1', MASTER_PORT='29613', PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29613', str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count('ok') == 2


def test_kl_divergence_follows_the_reference_formula():
    """KMP.KL_divergence (KMP.py:196-212, unused upstream) is host arithmetic: checked against scipy's densities."""
    from scipy.stats import multivariate_normal
    from kmp.model import KMP
    rng = np.random.default_rng(3)
    N, O = 17, 2
    k = KMP.__new__(KMP)
    k.N = N
    xi, xr = rng.normal(size=(O, N)), rng.normal(size=(O, N))
    B1, B2 = rng.normal(size=(N, O, O)), rng.normal(size=(N, O, O))
    sg = np.transpose(np.einsum('nab,ncb->nac', B1, B1) + 0.2 * np.eye(O), (1, 2, 0))
    sr = np.transpose(np.einsum('nab,ncb->nac', B2, B2) + 0.2 * np.eye(O), (1, 2, 0))
    ref = []
    for i in range(N):
        p = multivariate_normal(xi[:, i], sg[:, :, i]).pdf(xi[:, i])
        q = multivariate_normal(xr[:, i], sr[:, :, i]).pdf(xi[:, i])
        ref.append(q * np.log(q / p))
    ref = np.array(ref); ref /= np.linalg.norm(ref)
    assert abs(k.KL_divergence(xi, sg, xr, sr) - np.mean(ref)) < 1e-13


def test_headless_demos_import_like_main_py():
    """main.py:12 `from kmp import demo1, demo2, demo3` resolves without a display or a device; unknown names still fail."""
    import kmp
    from kmp import demo1, demo2, demo3
    from kmp.demos import demo1 as d1
    assert d1 is demo1 and {c.__name__ for c in (demo1, demo2, demo3)} == {'demo1', 'demo2', 'demo3'}
    with pytest.raises(AttributeError):
        kmp.demo4


@pytest.mark.parametrize('warps', [4, 8])
@pytest.mark.parametrize('O', [1, 2])
def test_predict_stage_schedule_covers_every_stage_exactly_once(O, warps):
    """Host side of the table-driven predict kernel (predict_tab.cu): for every (O, N) the stage lists of the MMA warps must be a
    partition of the lower-triangular contraction V = Linv k*^T (KMP.py:187-190 restated as chunks of 16 rows x blocks of 32
    columns per component block) - every stage once, the stages of a chunk consecutive in ONE warp's list and in k order, the
    chunk-end flag on the last of them, the V_1 flag where a chunk reaches the second component block."""
    from kmp_demos_b200 import _lib
    lib = _lib.load()
    row = (ctypes.c_int * 776)()
    assert lib.kmpx_debug_predict_schedule(3, 10, warps, row, 776) == -1
    assert lib.kmpx_debug_predict_schedule(O, 10, 6, row, 776) == -3
    assert lib.kmpx_debug_predict_schedule(O, 10, warps, row, 100) == -4
    cap = 768 // 2 // warps                                   # stages per warp the table holds
    for N in (1, 2, 7, 8, 9, 16, 31, 32, 33, 100, 201, 202, 204, 208, 224, 250, 300):
        rc = lib.kmpx_debug_predict_schedule(O, N, warps, row, 776)
        assert rc in (0, 776)
        Ns = (N + 1) & ~1
        n = O * Ns
        expect = []                                            # (r0, a, k offset, k-steps, first column of the block)
        for r0 in range(0, n, 16):
            for a in range(O):
                jm = min(max(r0 + 16 - a * Ns, 0), N)
                for kb in range((jm + 31) // 32):
                    kv = jm - 32 * kb
                    expect.append((r0, a, 32 * kb, 8 if kv >= 32 else 2 * ((kv + 7) // 8), a * Ns))
        if rc == 0:
            assert len(expect) > cap                          # refused only when it cannot fit
            continue
        got, load = [], []
        for w in range(warps):
            cnt = row[w]
            assert 0 <= cnt <= cap
            stages = []
            for i in range(cnt):
                d0, d1 = row[8 + 2 * (w * cap + i)], row[8 + 2 * (w * cap + i) + 1]
                stages.append(dict(r0=d0 & 0xfff, a=(d0 >> 12) & 1, end=(d0 >> 13) & 1, v1=(d0 >> 14) & 1, nst=d0 >> 16,
                                   k=d1 & 0xffff, col0=d1 >> 16))
            # chunks are runs of equal r0; inside a run: component 0 before 1, k ascending, the end flag on the last stage only
            i = 0
            seen = set()
            while i < cnt:
                j = i
                while j + 1 < cnt and stages[j + 1]['r0'] == stages[i]['r0']:
                    j += 1
                run = stages[i:j + 1]
                assert stages[i]['r0'] not in seen           # one run per chunk
                seen.add(stages[i]['r0'])
                assert [(s['a'], s['k']) for s in run] == sorted((s['a'], s['k']) for s in run)
                assert [s['end'] for s in run] == [0] * (len(run) - 1) + [1]
                has_v1 = int(any(s['a'] == 1 for s in run))
                assert all(s['v1'] == has_v1 for s in run)
                i = j + 1
            got += [(s['r0'], s['a'], s['k'], s['nst'], s['col0']) for s in stages]
            load.append(sum(32 * s['nst'] + 120 for s in stages))
        assert sorted(got) == sorted(expect), (O, N, warps)
        if len(expect) >= 8 * warps:                          # the deal is balanced (cycle model of the kernel: 32 per k-step + hand-over)
            assert max(load) <= 1.25 * (sum(load) / warps), (O, N, warps, load)
