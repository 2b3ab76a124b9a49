"""``GaussianMixtureModel`` - drop-in for the reference's ``kmp.mixture.GaussianMixtureModel``
(kmp/mixture/GaussianMixtureModel.py), which wraps ``sklearn.mixture.GaussianMixture`` and adds GMR.

  fit      GaussianMixtureModel.py:44-57 -> sklearn GaussianMixture(n_components, reg_covar, random_state=420):
           k-means++ / Lloyd initialisation and the EM loop run on the device (kmpx_kpp_candidates,
           kmpx_kmeans_assign, kmpx_gmm_label_stats, kmpx_gmm_em_pass, kmpx_gmm_finalize).  The host keeps what
           is inherently sequential in sklearn: the RandomState(420) stream, the cumulative-sum sampling of
           k-means++ (sklearn/cluster/_kmeans.py:227-231) and the convergence tests.
  predict  GaussianMixtureModel.py:59-107 (GMR) -> kmpx_gmr.
  bic      GaussianMixtureModel.py:109-122 -> one E-step reduction.

When the samples are partitioned across GPUs (``EMEngine(..., group=...)``) the only exchange is one
all-reduce (sum) of the packed sufficient statistics per EM iteration (NCCL over NVLink via torch.distributed).
"""
from __future__ import annotations

import logging
import math
from typing import Tuple

import numpy as np

from . import _lib

logging.basicConfig(
    level=logging.INFO,
    format='%(asctime)s.%(msecs)03d %(levelname)s %(message)s',
    datefmt='%Y-%m-%d,%H:%M:%S'
)

REALMIN = np.finfo(np.float64).tiny


def stats_len(K, d):
    return K * (1 + d + d * (d + 1) // 2) + 1


class EMEngine:
    """Device-side EM for a full-covariance mixture on one partition of the samples.

    X: (n_local, d) float64 CUDA tensor.  ``group``: torch.distributed process group over which the
    sufficient statistics are summed (None = single partition).  ``n_total``: global sample count."""

    def __init__(self, X, K, reg, shift=None, group=None, n_total=None):
        torch = _lib.require_cuda()
        self.torch = torch
        lib = _lib.load()
        self.X = X.contiguous()
        self.n, self.d = self.X.shape
        self.K, self.reg, self.group = int(K), float(reg), group
        self.n_total = float(n_total if n_total is not None else self.n)
        dev = X.device
        f64 = torch.float64
        K, d = self.K, self.d
        self.shift = torch.zeros(d, dtype=f64, device=dev) if shift is None else shift.to(dev, f64).contiguous()
        self.stats = torch.zeros(stats_len(K, d), dtype=f64, device=dev)
        self.ws_bytes = int(lib.kmpx_gmm_workspace_bytes(int(min(self.n, 2 ** 31 - 1)), d, K))
        self.ws = torch.empty(self.ws_bytes // 8 + 1, dtype=f64, device=dev)
        self.weights = torch.empty(K, dtype=f64, device=dev)
        self.log_weights = torch.empty(K, dtype=f64, device=dev)
        self.means = torch.empty((K, d), dtype=f64, device=dev)
        self.covs = torch.empty((K, d, d), dtype=f64, device=dev)
        self.prec_chol = torch.empty((K, d, d), dtype=f64, device=dev)
        self.log_det = torch.empty(K, dtype=f64, device=dev)
        self.info = torch.zeros(K, dtype=torch.int32, device=dev)
        self.allreduce_count = 0
        self.allreduce_events = None      # set to a list to collect (start, end) CUDA events around every all-reduce

    def _allreduce(self):
        if self.group is not None:
            import torch.distributed as dist
            if self.allreduce_events is not None:
                e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
                e0.record()
                dist.all_reduce(self.stats, op=dist.ReduceOp.SUM, group=self.group)
                e1.record()
                self.allreduce_events.append((e0, e1))
            else:
                dist.all_reduce(self.stats, op=dist.ReduceOp.SUM, group=self.group)
            self.allreduce_count += 1

    def _finalize(self, first):
        _lib.call('kmpx_gmm_finalize', _lib.ptr(self.stats), self.n_total, int(first), self.d, self.K, self.reg,
                  _lib.ptr(self.shift), _lib.ptr(self.weights), _lib.ptr(self.log_weights), _lib.ptr(self.means),
                  _lib.ptr(self.covs), _lib.ptr(self.prec_chol), _lib.ptr(self.log_det), _lib.ptr(self.info), _lib.stream())

    def init_from_labels(self, labels):
        """One-hot responsibilities -> first M-step (sklearn/mixture/_base.py:119-128, _gaussian_mixture.py:849-862)."""
        _lib.call('kmpx_gmm_label_stats', _lib.ptr(self.X), self.n, self.d, self.K, _lib.ptr(self.shift),
                  _lib.ptr(labels), _lib.ptr(self.stats), _lib.ptr(self.ws), self.ws_bytes, _lib.stream())
        self._allreduce()
        self._finalize(first=True)

    def set_params(self, weights, means, covs):
        """Start EM from given parameters (host arrays (K,), (K,d), (K,d,d)): the precision factors are
        derived on the device from the equivalent statistics."""
        torch = self.torch
        K, d = self.K, self.d
        w = np.asarray(weights, float)
        m = np.asarray(means, float) - self.shift.cpu().numpy()[None, :]
        c = np.asarray(covs, float)
        F = 1 + d + d * (d + 1) // 2
        st = np.zeros(stats_len(K, d))
        iu = np.triu_indices(d)
        for k in range(K):
            nk = w[k] * self.n_total
            st[k * F] = nk
            st[k * F + 1:k * F + 1 + d] = nk * m[k]
            st[k * F + 1 + d:(k + 1) * F] = nk * ((c[k] - self.reg * np.eye(d)) + np.outer(m[k], m[k]))[iu]
        self.stats.copy_(torch.as_tensor(st, device=self.stats.device))
        self._finalize(first=True)

    def em_pass(self, labels=None, lse=None, update=True):
        """One fused E+M pass.  Returns nothing; ``self.stats[-1]`` holds the (global) log-likelihood sum of
        the parameters the pass started from.  update=False leaves the parameters untouched (score / predict)."""
        _lib.call('kmpx_gmm_em_pass', _lib.ptr(self.X), self.n, self.d, self.K, _lib.ptr(self.shift),
                  _lib.ptr(self.log_weights), _lib.ptr(self.means), _lib.ptr(self.prec_chol), _lib.ptr(self.log_det),
                  _lib.ptr(self.stats), _lib.ptr(labels), _lib.ptr(lse), _lib.ptr(self.ws), self.ws_bytes, _lib.stream())
        self._allreduce()
        if update:
            self._finalize(first=False)

    def run(self, tol=1e-3, max_iter=100, fixed_iters=None):
        """sklearn's loop (sklearn/mixture/_base.py:265-278): stop when |delta mean log-lik| < tol.
        fixed_iters runs exactly that many passes without reading anything back (benchmarks)."""
        if fixed_iters is not None:
            for _ in range(fixed_iters):
                self.em_pass()
            return fixed_iters, float(self.stats[-1].item()) / self.n_total, False
        lower = -np.inf
        converged = False
        n_iter = 0
        for it in range(1, max_iter + 1):
            prev = lower
            self.em_pass()
            lower = float(self.stats[-1].item()) / self.n_total
            n_iter = it
            if abs(lower - prev) < tol:
                converged = True
                break
        if int(self.info.abs().max().item()) != 0:
            raise ValueError("Fitting the mixture model failed because some components have ill-defined empirical "
                             "covariance (for instance caused by singleton or collapsed samples). Try to decrease the "
                             "number of components, increase reg_covar, or scale the input data.")
        return n_iter, lower, converged


def kmeans_init_labels(X, K, rs, max_iter=300, tol=1e-4):
    """``KMeans(n_clusters=K, n_init=1, random_state=rs).fit(X).labels_`` with the sample-sized work on the device.
    X: (n,d) CUDA tensor (not modified).  Returns int32 CUDA labels, the Lloyd iteration count and the
    data mean (CUDA, (d,)) that the statistics pass produced on the way."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = X.device
    f64, i32 = torch.float64, torch.int32
    n, d = X.shape
    st = _lib.stream()
    ws_bytes = max(int(lib.kmpx_gmm_workspace_bytes(int(n), d, max(K, 1))), 148 * 4 * 8 * 8)
    ws = torch.empty(ws_bytes // 8 + 1, dtype=f64, device=dev)
    F = 1 + d + d * (d + 1) // 2
    zeros_lab = torch.zeros(n, dtype=i32, device=dev)
    # data mean and per-feature variance from one statistics pass (K=1)
    st1 = torch.zeros(F + 1, dtype=f64, device=dev)
    _lib.call('kmpx_gmm_label_stats', _lib.ptr(X), n, d, 1, None, _lib.ptr(zeros_lab), _lib.ptr(st1), _lib.ptr(ws), ws_bytes, st)
    s1 = st1.cpu().numpy()
    mean = s1[1:1 + d] / n
    Xc = X.clone()
    x_sq = torch.empty(n, dtype=f64, device=dev)
    mean_dev = torch.as_tensor(mean, device=dev)
    _lib.call('kmpx_rows_center_sqnorm', _lib.ptr(Xc), n, d, _lib.ptr(mean_dev), _lib.ptr(x_sq), st)
    _lib.call('kmpx_gmm_label_stats', _lib.ptr(Xc), n, d, 1, None, _lib.ptr(zeros_lab), _lib.ptr(st1), _lib.ptr(ws), ws_bytes, st)
    s1 = st1.cpu().numpy()
    diag_idx = np.cumsum([0] + [d - i for i in range(d - 1)])
    var = s1[1 + d + diag_idx] / n - (s1[1:1 + d] / n) ** 2
    tol_abs = float(np.mean(var) * tol)
    # ---- k-means++ (sklearn/cluster/_kmeans.py:180-278)
    n_trials = 2 + int(np.log(K))
    w = np.ones(n)
    center_ids = np.full(K, -1, dtype=np.int64)
    center_ids[0] = rs.choice(n, p=w / w.sum())
    cand = torch.as_tensor(np.array([center_ids[0]], dtype=np.int32), device=dev)
    dist = torch.empty((max(n_trials, 1), n), dtype=f64, device=dev)
    pot_dev = torch.empty(8, dtype=f64, device=dev)
    _lib.call('kmpx_kpp_candidates', _lib.ptr(Xc), _lib.ptr(x_sq), n, d, _lib.ptr(cand), 1, None, _lib.ptr(dist),
              _lib.ptr(pot_dev), _lib.ptr(ws), ws_bytes, st)
    closest = dist[0].clone()
    pot = float(pot_dev[0].item())
    for c in range(1, K):
        rand_vals = rs.uniform(size=n_trials) * pot
        cum = np.cumsum(closest.cpu().numpy())
        cids = np.searchsorted(cum, rand_vals)
        np.clip(cids, None, n - 1, out=cids)
        cand = torch.as_tensor(cids.astype(np.int32), device=dev)
        _lib.call('kmpx_kpp_candidates', _lib.ptr(Xc), _lib.ptr(x_sq), n, d, _lib.ptr(cand), n_trials, _lib.ptr(closest),
                  _lib.ptr(dist), _lib.ptr(pot_dev), _lib.ptr(ws), ws_bytes, st)
        pots = pot_dev[:n_trials].cpu().numpy()
        best = int(np.argmin(pots))
        pot = float(pots[best])
        closest = dist[best].clone()
        center_ids[c] = cids[best]
    centers = Xc[torch.as_tensor(center_ids, device=dev)].contiguous()          # (K,d) gather of the chosen rows
    # ---- Lloyd (sklearn/cluster/_kmeans.py:630-758)
    labels = torch.full((n,), -1, dtype=i32, device=dev)
    labels_new = torch.empty(n, dtype=i32, device=dev)
    n_changed = torch.zeros(1, dtype=i32, device=dev)
    stats = torch.zeros(stats_len(K, d), dtype=f64, device=dev)
    strict = False
    n_iter = 0
    for it in range(max_iter):
        n_changed.zero_()
        _lib.call('kmpx_kmeans_assign', _lib.ptr(Xc), n, d, K, _lib.ptr(centers), _lib.ptr(labels), _lib.ptr(labels_new),
                  _lib.ptr(n_changed), st)
        _lib.call('kmpx_gmm_label_stats', _lib.ptr(Xc), n, d, K, None, _lib.ptr(labels_new), _lib.ptr(stats), _lib.ptr(ws),
                  ws_bytes, st)
        sh = stats.cpu().numpy()[:-1].reshape(K, F)
        cnt = sh[:, 0].copy()
        sums = sh[:, 1:1 + d].copy()
        old = centers.cpu().numpy()
        if np.any(cnt == 0):
            sums, cnt = _relocate_empty(Xc, labels_new, old, sums, cnt)
        new = np.empty_like(old)
        amax = int(np.argmax(cnt))
        for j in range(K):
            new[j] = sums[j] * (1.0 / cnt[j]) if cnt[j] > 0 else np.nan
        for j in range(K):
            if not cnt[j] > 0:
                new[j] = new[amax]
        shift_tot = float(((new - old) ** 2).sum())
        centers = torch.as_tensor(new, device=dev)
        labels, labels_new = labels_new, labels
        n_iter = it + 1
        if int(n_changed.item()) == 0:
            strict = True
            break
        if shift_tot <= tol_abs:
            break
    if not strict:
        _lib.call('kmpx_kmeans_assign', _lib.ptr(Xc), n, d, K, _lib.ptr(centers), None, _lib.ptr(labels_new), None, st)
        labels = labels_new
    return labels, n_iter, mean_dev


def _relocate_empty(Xc, labels, centers_old, sums, cnt):
    """sklearn's _relocate_empty_clusters_dense (_k_means_common.pyx:167-211): an empty cluster takes the sample
    farthest from its own centre.  Rare; the distances are one fused device expression."""
    torch = _lib.require_cuda()
    empty = np.where(cnt == 0)[0]
    c_old = torch.as_tensor(centers_old, device=Xc.device)
    d2 = ((Xc - c_old[labels.long()]) ** 2).sum(dim=1)
    if float(d2.max().item()) == 0.0:
        return sums, cnt
    far = torch.topk(d2, len(empty)).indices.cpu().numpy()
    lab = labels.cpu().numpy()
    Xh = Xc[torch.as_tensor(far, device=Xc.device)].cpu().numpy()
    for e, f, x in zip(empty, far, Xh):
        old = lab[f]
        sums[old] -= x
        sums[e] = x
        cnt[e] = 1.0
        cnt[old] -= 1.0
    return sums, cnt


class _FittedMixture:
    """What the reference exposes as ``GaussianMixtureModel.model`` (an sklearn estimator): the fitted attributes
    and the three methods the demos and the wrapper call."""

    def __init__(self, owner):
        self._o = owner
        self.n_components = owner.n_components
        self.reg_covar = owner.diag_reg_factor
        self.random_state = 420
        self.tol, self.max_iter = 1e-3, 100

    def _engine(self, X):
        torch = _lib.require_cuda()
        Xd = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64), device='cuda')
        eng = EMEngine(Xd, self.n_components, self.reg_covar)
        eng.set_params(self.weights_, self.means_, self.covariances_)
        return eng, Xd

    def score_samples(self, X):
        torch = _lib.require_cuda()
        eng, Xd = self._engine(X)
        lse = torch.empty(Xd.shape[0], dtype=torch.float64, device='cuda')
        eng.em_pass(lse=lse, update=False)
        return lse.cpu().numpy()

    def score(self, X, y=None):
        eng, Xd = self._engine(X)
        eng.em_pass(update=False)
        return float(eng.stats[-1].item()) / Xd.shape[0]

    def predict(self, X):
        torch = _lib.require_cuda()
        eng, Xd = self._engine(X)
        labels = torch.empty(Xd.shape[0], dtype=torch.int32, device='cuda')
        eng.em_pass(labels=labels, update=False)
        return labels.cpu().numpy().astype(np.int64)

    def _n_parameters(self):
        K, d = self.means_.shape
        return int(K * d * (d + 1) / 2.0 + d * K + K - 1)

    def bic(self, X):
        return -2 * self.score(X) * X.shape[0] + self._n_parameters() * math.log(X.shape[0])


class GaussianMixtureModel:
    """Gaussian mixture model + Gaussian mixture regression (B200 implementation).

    Parameters (identical to the reference, GaussianMixtureModel.py:33-36)
    ----------
    n_components : int, default=10
    n_demos : int, default=5          stored, unused in compute (as in the reference)
    diag_reg_factor : float, default=1e-4
    """

    def __init__(self, n_components: int = 10, n_demos: int = 5, diag_reg_factor: float = 1e-4) -> None:
        self.n_components = n_components
        self.n_demos = n_demos
        self.diag_reg_factor = diag_reg_factor
        self.model = _FittedMixture(self)
        self.logger = logging.getLogger(__name__)

    def fit(self, data: np.ndarray) -> None:
        """data: (n_samples, n_features).  Same result as sklearn's GaussianMixture(random_state=420).fit."""
        torch = _lib.require_cuda()
        data = np.ascontiguousarray(data, dtype=np.float64)
        n, d = data.shape
        K = self.n_components
        if n < K:
            raise ValueError(f"Expected n_samples >= n_components but got n_components = {K}, n_samples = {n}")
        if not 1 <= d <= 8 or not 1 <= K <= 128:
            raise ValueError(f"the device EM kernels cover 1 <= n_features <= 8 and 1 <= n_components <= 128 "
                             f"(got n_features = {d}, n_components = {K})")
        self.n_features = d
        X = torch.as_tensor(data, device='cuda')
        rs = np.random.RandomState(420)
        labels, _, shift = kmeans_init_labels(X, K, rs)
        eng = EMEngine(X, K, self.diag_reg_factor, shift=shift)
        eng.init_from_labels(labels)
        n_iter, lower, converged = eng.run(tol=1e-3, max_iter=100)
        m = self.model
        m.weights_ = eng.weights.cpu().numpy()
        m.means_ = eng.means.cpu().numpy()
        m.covariances_ = eng.covs.cpu().numpy()
        m.precisions_cholesky_ = eng.prec_chol.cpu().numpy()
        m.n_iter_, m.lower_bound_, m.converged_ = n_iter, lower, converged
        self.logger.info("GMM fit done.")
        self.priors = m.weights_
        self.means = m.means_.T
        self.covariances = np.transpose(m.covariances_, (1, 2, 0))

    def predict(self, data: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """GMR: data (I, M) -> means (O, M), covariances (O, O, M)."""
        torch = _lib.require_cuda()
        data = np.ascontiguousarray(data, dtype=np.float64)
        I, M = data.shape
        O = self.n_features - I
        dev = torch.device('cuda')
        f64 = torch.float64
        pri = torch.as_tensor(np.ascontiguousarray(self.priors, dtype=np.float64), device=dev)
        mea = torch.as_tensor(np.ascontiguousarray(self.means, dtype=np.float64), device=dev)
        cov = torch.as_tensor(np.ascontiguousarray(self.covariances, dtype=np.float64), device=dev)
        x = torch.as_tensor(data, device=dev)
        mu = torch.empty((O, M), dtype=f64, device=dev)
        sg = torch.empty((O, O, M), dtype=f64, device=dev)
        _lib.call('kmpx_gmr', _lib.ptr(pri), _lib.ptr(mea), _lib.ptr(cov), self.n_components, I, O, _lib.ptr(x), M,
                  float(self.diag_reg_factor), _lib.ptr(mu), _lib.ptr(sg), _lib.stream())
        self.logger.info("GMR done.")
        return mu.cpu().numpy(), sg.cpu().numpy()

    def bic(self, data: np.ndarray) -> float:
        return self.model.bic(np.asarray(data, dtype=np.float64))
