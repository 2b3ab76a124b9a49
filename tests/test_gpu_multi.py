"""Multi-GPU path on real devices (needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
EM on row-partitioned samples with one NCCL all-reduce of the sufficient statistics per iteration must equal the
single-GPU EM within 1e-12; the sharded batched KMP run must equal the unsharded one
bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["KMP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from kmp_demos_b200 import dist as kd, workloads as wl
from kmp_demos_b200.mixture import EMEngine
rank, world, local = kd.init_process_group("nccl")
dev = torch.device("cuda", local)
n, K, d = 200000, 16, 7
X, means, _ = wl.cfg5_samples(n, K=K, d=d, seed=5)
lab = np.argmin(((X[:, None, :] - means[None]) ** 2).sum(-1), axis=1).astype(np.int32)
shift = torch.as_tensor(X.mean(axis=0), device=dev)
lo, hi = kd.shard_range(n, rank, world)
eng = EMEngine(torch.as_tensor(X[lo:hi], device=dev), K, 1e-6, shift=shift, group=dist.group.WORLD, n_total=n)
eng.init_from_labels(torch.as_tensor(lab[lo:hi], device=dev))
n_iter, lower, conv = eng.run(tol=1e-3, max_iter=100, check_every=1)      # one all-reduce per executed pass
out = dict(n_iter=n_iter, lower=lower, allreduces=eng.allreduce_count)
if rank == 0:
    ref = EMEngine(torch.as_tensor(X, device=dev), K, 1e-6, shift=shift)
    ref.init_from_labels(torch.as_tensor(lab, device=dev))
    n2, lower2, _ = ref.run(tol=1e-3, max_iter=100, check_every=1)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    out.update(n_iter_ref=n2, lower_ref=lower2, err_means=rel(eng.means, ref.means), err_covs=rel(eng.covs, ref.covs),
               err_weights=rel(eng.weights, ref.weights))
    print("RESULT " + json.dumps(out))
dist.barrier()
dist.destroy_process_group()
'''


def test_partitioned_em_with_nccl_allreduce_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    env = dict(os.environ, KMP_ROOT=ROOT, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29631', str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')][0]
    out = json.loads(line[7:])
    assert out['n_iter'] == out['n_iter_ref'] and out['allreduces'] == out['n_iter'] + 1
    assert abs(out['lower'] - out['lower_ref']) < 1e-12 * abs(out['lower_ref'])
    assert out['err_means'] < 1e-12 and out['err_covs'] < 1e-12 and out['err_weights'] < 1e-12


_WORKER_FIT = r'''
import os, sys, json
sys.path.insert(0, os.environ["KMP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from kmp_demos_b200 import dist as kd, workloads as wl
from kmp_demos_b200.mixture import GaussianMixtureModel
rank, world, local = kd.init_process_group("nccl")
dev = torch.device("cuda", local)
z = np.load(os.path.join(os.environ["KMP_ROOT"], "tests", "golden", "gmm_synth.npz"))
n, K = int(z["n"]), int(z["K"])
X, _, _ = wl.cfg5_samples(n, K=K, d=7, seed=int(z["seed"]))
lo, hi = kd.shard_range(n, rank, world)
g = GaussianMixtureModel(n_components=K, diag_reg_factor=1e-6)
g.logger.disabled = True
g.fit_device(torch.as_tensor(X[lo:hi], device=dev), group=dist.group.WORLD, n_total=n, n_before=lo)
rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
out = dict(n_iter=int(g.model.n_iter_), n_iter_ref=int(z["n_iter"]), err_priors=rel(g.priors, z["priors"]),
           err_means=rel(g.means, z["means"]), err_covs=rel(g.covariances, z["covs"]))
if rank == 0:
    print("RESULT " + json.dumps(out))
dist.barrier()
dist.destroy_process_group()
'''


def test_row_partitioned_fit_with_distributed_kmeans_init_equals_sklearn(tmp_path):
    """GaussianMixtureModel.fit_device on rows split over 2 GPUs (k-means++ / Lloyd / EM exchange packed sums only)
    reproduces the reference wrapper's sklearn fit of the whole data set: same EM iteration count, parameters
    1e-9 / 1e-7."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'worker_fit.py'
    script.write_text(_WORKER_FIT)
    env = dict(os.environ, KMP_ROOT=ROOT, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29635', str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')][0][7:])
    assert out['n_iter'] == out['n_iter_ref']
    assert out['err_priors'] < 1e-9 and out['err_means'] < 1e-9 and out['err_covs'] < 1e-7


_WORKER_F3 = r'''
import os, sys, json
sys.path.insert(0, os.environ["KMP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from kmp_demos_b200 import dist as kd, workloads as wl
from kmp.model import KMP
rank, world, local = kd.init_process_group("nccl")
t, mu, sig = wl.cfg4_problem(N=768, seed=0)                       # n = 1536: the large-system kernels
q = np.linspace(0.0, 2.0, 1001).reshape(1, -1)
c = wl.CFG1
k = KMP(l=c["lam"], alpha=c["alpha"], sigma_f=c["sigma_f"], time_driven_kernel=False)
if rank == 0:
    k.fit(t, mu, sig)                                             # only one GPU factorises
k.broadcast_fit(src=0)
xi, sg = k.predict_sharded(q)
if rank == 0:
    x1, s1 = k.predict(q)
    print("RESULT " + json.dumps(dict(same_mean=bool(np.array_equal(xi, x1)), same_cov=bool(np.array_equal(sg, s1)))))
dist.barrier()
dist.destroy_process_group()
'''


def test_query_sharded_predict_of_one_large_system(tmp_path):
    """rank 0 factorises, the factor is broadcast over NCCL, every rank predicts its slice of the
    queries; the gathered result equals the single-GPU prediction bit for bit."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'worker_f3.py'
    script.write_text(_WORKER_F3)
    env = dict(os.environ, KMP_ROOT=ROOT, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29633', str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')][0][7:])
    assert out['same_mean'] and out['same_cov']
