"""``GaussianMixtureModel`` - drop-in for the reference's ``kmp.mixture.GaussianMixtureModel``
(kmp/mixture/GaussianMixtureModel.py), which wraps ``sklearn.mixture.GaussianMixture`` and adds GMR.

  fit      GaussianMixtureModel.py:44-57 -> sklearn GaussianMixture(n_components, reg_covar, random_state=420):
           k-means++ sampling and greedy choice (kmpx_kpp_sample / _candidates_rows / _choose), Lloyd (kmpx_kmeans_assign,
           kmpx_gmm_label_sums, kmpx_lloyd_update) and the EM loop with sklearn's stopping rule (kmpx_gmm_label_stats,
           kmpx_gmm_em_pass, kmpx_gmm_finalize_check) run on the device.  The host keeps the RandomState(420) stream and
           reads one small flag block per centre / per batch of iterations.
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

    def run(self, tol=1e-3, max_iter=100, fixed_iters=None, check_every=None):
        """sklearn's loop (sklearn/mixture/_base.py:265-278): stop when |delta mean log-lik| < tol.  The comparison runs
        on the device (kmpx_gmm_finalize_check): passes are enqueued `check_every` at a time and the host reads the
        4-double loop state once per batch (a pass enqueued after convergence leaves the parameters untouched).
        fixed_iters runs exactly that many passes without reading anything back (benchmarks)."""
        torch = self.torch
        if fixed_iters is not None:
            for _ in range(fixed_iters):
                self.em_pass()
            return fixed_iters, float(self.stats[-1].item()) / self.n_total, False
        if check_every is None:
            # one read per pass costs ~20 us: irrelevant next to a pass over millions of rows, dominant for a few hundred
            check_every = 1 if self.n * self.K >= 20_000_000 else 8
        state = torch.tensor([0.0, 0.0, -float('inf'), 0.0], dtype=torch.float64, device=self.X.device)
        st = None
        while True:
            for _ in range(check_every):
                _lib.call('kmpx_gmm_em_pass', _lib.ptr(self.X), self.n, self.d, self.K, _lib.ptr(self.shift),
                          _lib.ptr(self.log_weights), _lib.ptr(self.means), _lib.ptr(self.prec_chol), _lib.ptr(self.log_det),
                          _lib.ptr(self.stats), None, None, _lib.ptr(self.ws), self.ws_bytes, _lib.stream())
                self._allreduce()
                _lib.call('kmpx_gmm_finalize_check', _lib.ptr(self.stats), self.n_total, self.d, self.K, self.reg,
                          _lib.ptr(self.shift), _lib.ptr(self.weights), _lib.ptr(self.log_weights), _lib.ptr(self.means),
                          _lib.ptr(self.covs), _lib.ptr(self.prec_chol), _lib.ptr(self.log_det), _lib.ptr(self.info),
                          float(tol), int(max_iter), _lib.ptr(state), _lib.stream())
            st = state.cpu().numpy()
            if st[0] != 0.0:
                break
        n_iter, lower, converged = int(st[1]), float(st[2]), bool(st[3] != 0.0)
        if int(self.info.abs().max().item()) != 0:
            raise ValueError("Fitting the mixture model failed because some components have ill-defined empirical "
                             "covariance (for instance caused by singleton or collapsed samples). Try to decrease the "
                             "number of components, increase reg_covar, or scale the input data.")
        return n_iter, lower, converged


class _Ranks:
    """The optional process group behind a row-partitioned fit: rank / world and the three exchanges the k-means
    initialisation needs (sum, gather of equal-sized blocks, gather of ragged row vectors to the host)."""

    def __init__(self, group):
        self.group = group
        if group is None:
            self.rank, self.world = 0, 1
        else:
            import torch.distributed as dist
            self.dist = dist
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def sum_(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def gather(self, t):
        """(world, *t.shape) tensor holding every rank's t."""
        if self.world == 1:
            return t.unsqueeze(0)
        out = t.new_empty((self.world,) + tuple(t.shape))
        self.dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group)
        return out

    def gather_rows_host(self, t, counts, root_only=False):
        """Concatenation over the ranks of the leading-dimension slices t[:counts[rank]] as a numpy array (single rank:
        through a pinned staging buffer that is kept for the next call).  root_only: only rank 0 copies the gathered
        rows to the host and returns them (the others return None)."""
        if self.world == 1:
            import torch
            buf = getattr(self, '_pinned', None)
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = self._pinned = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            buf.copy_(t)
            return buf.numpy()
        per = max(counts)
        pad = t.new_zeros((per,) + tuple(t.shape[1:]))
        pad[:t.shape[0]] = t
        allr = self.gather(pad)
        if root_only and self.rank != 0:
            return None
        allr = allr.cpu().numpy()
        return np.concatenate([allr[r, :counts[r]] for r in range(self.world)])


def _first_center(rs, n_total):
    """``rs.choice(n_total, p=ones/n_total)`` (sklearn/cluster/_kmeans.py:228) without materialising the n_total-long
    probability vector: numpy draws ONE uniform u and returns searchsorted(cumsum(p)/cumsum(p)[-1], u, 'right'), i.e.
    floor(u * n_total) unless u lies within the rounding error of that cumulative sum (< 4 (n+2) eps) of a boundary -
    then the literal call is replayed from the saved generator state."""
    state = rs.get_state()
    u = rs.random_sample()
    pos = u * n_total
    idx = int(np.floor(pos))
    margin = min(pos - idx, idx + 1 - pos) / n_total
    if margin <= 4.0 * (n_total + 2) * np.finfo(float).eps or idx >= n_total:
        rs.set_state(state)
        w = np.ones(n_total)
        idx = int(rs.choice(n_total, p=w / w.sum()))
    return idx


def kmeans_init_labels(X, K, rs, max_iter=300, tol=1e-4, group=None, n_total=None, n_before=0, host_sampling=False,
                       resolve_on_host=None):
    """``KMeans(n_clusters=K, n_init=1, random_state=rs).fit(X).labels_`` on the device, optionally with the rows
    partitioned over the ranks of `group` (this rank holds rows n_before .. n_before + n of n_total; every rank passes
    an identical `rs`).  X: (n,d) CUDA tensor (not modified).  The host keeps the RandomState stream only; k-means++
    sampling, greedy choice, Lloyd updates and stopping rules are kernels, the host reads one small flag block per
    centre / per batch of Lloyd iterations.  A k-means++ draw the device cannot decide provably the way numpy's
    sequential cumsum does (kmpx_kpp_sample) is repeated on the host with numpy's own expressions; host_sampling=True
    forces that path for every centre (tests).  Such a draw lies within the rounding error of numpy's cumulative sum /
    BLAS potential (up to n u pot each), i.e. where scikit-learn's own choice depends on its BLAS build and thread
    count; the probability is ~4 n^2 u per draw (4 % at n = 10 M, 2e-5 at n = 200 k).  resolve_on_host=False keeps the
    device's (exact-arithmetic) choice there instead of paying ~30 ms of host work per draw; the default resolves on
    the host up to 2 M rows - where parity with the fixtures is checked - and not beyond.  Returns int32 CUDA labels of
    the local rows, the Lloyd iteration count, the data mean (CUDA, (d,)) and a dict of counters."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = X.device
    f64, i32, i64 = torch.float64, torch.int32, torch.int64
    n, d = X.shape
    rk = _Ranks(group)
    n_total = int(n_total if n_total is not None else n)
    if resolve_on_host is None:
        resolve_on_host = n_total <= 2_000_000
    counts = [n] if rk.world == 1 else [int(v) for v in rk.gather(torch.tensor([n], dtype=i64, device=dev)).view(-1).tolist()]
    st = _lib.stream()
    ws_bytes = max(int(lib.kmpx_gmm_workspace_bytes(int(n), d, max(K, 1))), 148 * 4 * 16 * 8,
                   int(lib.kmpx_kpp_sample_workspace_bytes(int(n))))
    ws = torch.empty(ws_bytes // 8 + 1, dtype=f64, device=dev)
    F = 1 + d + d * (d + 1) // 2
    zeros_lab = torch.zeros(n, dtype=i32, device=dev)
    # ---- data mean, centring, per-feature variance -> Lloyd tolerance (sklearn/cluster/_kmeans.py:1486-1496, 285-295)
    st1 = torch.zeros(F + 1, dtype=f64, device=dev)
    _lib.call('kmpx_gmm_label_stats', _lib.ptr(X), n, d, 1, None, _lib.ptr(zeros_lab), _lib.ptr(st1), _lib.ptr(ws), ws_bytes, st)
    rk.sum_(st1)
    mean_dev = (st1[1:1 + d] / n_total).contiguous()
    Xc = X.clone()
    x_sq = torch.empty(n, dtype=f64, device=dev)
    _lib.call('kmpx_rows_center_sqnorm', _lib.ptr(Xc), n, d, _lib.ptr(mean_dev), _lib.ptr(x_sq), st)
    _lib.call('kmpx_gmm_label_stats', _lib.ptr(Xc), n, d, 1, None, _lib.ptr(zeros_lab), _lib.ptr(st1), _lib.ptr(ws), ws_bytes, st)
    rk.sum_(st1)
    diag_idx = torch.as_tensor(np.cumsum([0] + [d - i for i in range(d - 1)]) + 1 + d, device=dev)
    var = st1[diag_idx] / n_total - (st1[1:1 + d] / n_total) ** 2
    tol_dev = (var.mean() * tol).reshape(1).contiguous()
    # ---- k-means++ (sklearn/cluster/_kmeans.py:180-278)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    T = 2 + int(np.log(K))
    gid0 = _first_center(rs, n_total)
    U = rs.uniform(size=(max(K - 1, 1), T)) if K > 1 else np.zeros((1, T))
    U_dev = torch.as_tensor(U, device=dev)
    cand = torch.full((8,), -1, dtype=i64, device=dev)
    flags = torch.zeros(9, dtype=i32, device=dev)           # [0:T] draws, [8] greedy choice
    rows = torch.zeros((8, d + 2), dtype=f64, device=dev)
    dist = torch.empty((T, n), dtype=f64, device=dev)
    closest = [torch.empty(n, dtype=f64, device=dev), torch.empty(n, dtype=f64, device=dev)]
    pot_l = torch.zeros((T, 2), dtype=f64, device=dev)
    pot_dev = torch.zeros(1, dtype=f64, device=dev)
    before = torch.zeros(2, dtype=f64, device=dev)
    centers = torch.zeros((K, d), dtype=f64, device=dev)
    center_ids = torch.full((K,), -1, dtype=i64, device=dev)
    best = torch.zeros(1, dtype=i32, device=dev)
    info = {'kpp_host_draws': 0, 'kpp_host_choices': 0, 'kpp_ambiguous_draws': 0, 'lloyd_relocations': 0}

    def local_of(gids):
        """global row indices -> this rank's local indices (-1: another rank owns the row)"""
        g = np.asarray(gids, dtype=np.int64)
        loc = np.where((g >= n_before) & (g < n_before + n), g - n_before, -1)
        out = np.full(8, -1, dtype=np.int64)
        out[:len(loc)] = loc
        return torch.as_tensor(out, device=dev)

    def evaluate(cand_t, t_cnt, cur):
        """candidate rows -> everybody; distances / per-rank potentials of this rank's rows"""
        _lib.call('kmpx_kpp_gather_rows', _lib.ptr(Xc), d, _lib.ptr(cand_t), t_cnt, n_before, _lib.ptr(rows), st)
        rk.sum_(rows)
        _lib.call('kmpx_kpp_candidates_rows', _lib.ptr(Xc), _lib.ptr(x_sq), n, d, _lib.ptr(rows), t_cnt,
                  _lib.ptr(closest[cur]) if cur is not None else None, _lib.ptr(dist), _lib.ptr(pot_l), _lib.ptr(ws), ws_bytes, st)
        return rk.gather(pot_l[:t_cnt].contiguous())

    def choose(pots, t_cnt, c, nxt, force=-1):
        _lib.call('kmpx_kpp_choose', _lib.ptr(pots), rk.world, t_cnt, rk.rank, _lib.ptr(rows), d, c, n_total, force,
                  _lib.ptr(dist), n, _lib.ptr(pot_dev), _lib.ptr(before), _lib.ptr(centers), _lib.ptr(center_ids),
                  _lib.ptr(best), _lib.ptr(flags[8:]), _lib.ptr(closest[nxt]), st)

    def host_choice(t_cnt):
        """np.argmin(distance_to_candidates @ ones) with numpy's own product (_kmeans.py:241-245)"""
        full = rk.gather_rows_host(dist[:t_cnt].t().contiguous(), counts, root_only=True)          # (n_total, T)
        b = int(np.argmin(full.T @ np.ones((n_total, 1)))) if rk.rank == 0 else 0
        if rk.world > 1:
            bd = torch.tensor([b], dtype=i64, device=dev)
            rk.dist.broadcast(bd, src=rk.dist.get_global_rank(rk.group, 0) if rk.group is not rk.dist.group.WORLD else 0, group=rk.group)
            b = int(bd.item())
        return b

    pots = evaluate(local_of([gid0]), 1, None)
    choose(pots, 1, 0, 0)
    cur = 0
    host_ones = host_cum = None
    for c in range(1, K):
        nxt = 1 - cur
        if not host_sampling:
            _lib.call('kmpx_kpp_sample', _lib.ptr(closest[cur]), n, _lib.ptr(U_dev[c - 1]), T, _lib.ptr(pot_dev), _lib.ptr(before),
                      n_before, n_total, int(rk.rank == rk.world - 1), _lib.ptr(cand), _lib.ptr(flags), _lib.ptr(ws), ws_bytes, st)
            pots = evaluate(cand, T, cur)                    # speculative: queued before the flags are read
            choose(pots, T, c, nxt)
            fl = rk.sum_(flags.clone()).cpu().numpy() if rk.world > 1 else flags.cpu().numpy()
            redraw = bool(fl[:T].any())
            info['kpp_ambiguous_draws'] += int(redraw)
            if not resolve_on_host:
                redraw = False
                fl[8] = 0
        else:
            redraw, fl = True, None
        if redraw:
            # the draw lies within numpy's own rounding uncertainty: numpy's expressions decide (_kmeans.py:227-233)
            info['kpp_host_draws'] += 1
            full = rk.gather_rows_host(closest[cur], counts, root_only=True)
            if rk.rank == 0:
                if host_ones is None:
                    host_ones, host_cum = np.ones(n_total), np.empty(n_total)
                rand_vals = U[c - 1] * (full @ host_ones)
                cids = np.searchsorted(np.cumsum(full, out=host_cum), rand_vals)
                np.clip(cids, None, n_total - 1, out=cids)
            else:
                cids = np.zeros(T, dtype=np.int64)
            if rk.world > 1:                                  # one rank does the host arithmetic, everybody gets the indices
                cd = torch.as_tensor(np.asarray(cids, dtype=np.int64), device=dev)
                rk.dist.broadcast(cd, src=rk.dist.get_global_rank(rk.group, 0) if rk.group is not rk.dist.group.WORLD else 0, group=rk.group)
                cids = cd.cpu().numpy()
            pots = evaluate(local_of(cids), T, cur)
            choose(pots, T, c, nxt)
            fl = rk.sum_(flags.clone()).cpu().numpy() if rk.world > 1 else flags.cpu().numpy()
        if fl[8]:
            info['kpp_host_choices'] += 1
            choose(pots, T, c, nxt, force=host_choice(T))
        cur = nxt
    # ---- Lloyd (sklearn/cluster/_kmeans.py:630-758)
    ev[1].record()
    labels = torch.full((n,), -1, dtype=i32, device=dev)
    labels_new = torch.empty(n, dtype=i32, device=dev)
    n_changed = torch.zeros(1, dtype=i32, device=dev)
    stats = torch.zeros(stats_len(K, d), dtype=f64, device=dev)
    state = torch.zeros(4, dtype=i32, device=dev)
    batch = 1 if n * K >= 20_000_000 else 8

    def iteration():
        nonlocal labels, labels_new
        _lib.call('kmpx_kmeans_assign', _lib.ptr(Xc), n, d, K, _lib.ptr(centers), _lib.ptr(labels), _lib.ptr(labels_new),
                  _lib.ptr(n_changed), st)
        _lib.call('kmpx_gmm_label_sums', _lib.ptr(Xc), n, d, K, None, _lib.ptr(labels_new), _lib.ptr(stats), _lib.ptr(ws),
                  ws_bytes, st)
        stats[-1:].copy_(n_changed)                           # travels with the sums through the one all-reduce
        rk.sum_(stats)
        _lib.call('kmpx_lloyd_update', _lib.ptr(stats), K, d, _lib.ptr(centers), _lib.ptr(tol_dev), _lib.ptr(stats[-1:]),
                  _lib.ptr(n_changed), int(max_iter), _lib.ptr(state), st)
        labels, labels_new = labels_new, labels

    while True:
        for _ in range(batch):
            iteration()
        s_host = state.cpu().numpy()
        if s_host[3]:
            # an empty cluster (rare): this iteration's update is done on the host with sklearn's relocation rule.
            # The surplus iterations of the batch were no-ops, so labels / stats are those of the flagged iteration,
            # except for the label-change count, which is recovered from one more assignment against the kept labels
            info['lloyd_relocations'] += 1
            _lib.call('kmpx_kmeans_assign', _lib.ptr(Xc), n, d, K, _lib.ptr(centers), None, _lib.ptr(labels_new), None, st)
            _lib.call('kmpx_gmm_label_sums', _lib.ptr(Xc), n, d, K, None, _lib.ptr(labels_new), _lib.ptr(stats), _lib.ptr(ws),
                      ws_bytes, st)
            stats[-1:].zero_()
            rk.sum_(stats)
            sh = stats.cpu().numpy()[:-1].reshape(K, F)
            cnt, sums = sh[:, 0].copy(), sh[:, 1:1 + d].copy()
            old = centers.cpu().numpy()
            sums, cnt = _relocate_empty(Xc, labels_new, old, sums, cnt, rk, n_before)
            new = np.empty_like(old)
            amax = int(np.argmax(cnt))
            for j in range(K):
                new[j] = sums[j] * (1.0 / cnt[j]) if cnt[j] > 0 else np.nan
            for j in range(K):
                if not cnt[j] > 0:
                    new[j] = new[amax]
            centers.copy_(torch.as_tensor(new, device=dev))
            it_done = int(s_host[1]) + 1
            shift_tot = float(((new - old) ** 2).sum())
            finished = shift_tot <= float(tol_dev.item()) or it_done >= max_iter
            state.copy_(torch.tensor([int(finished), it_done, 0, 0], dtype=i32, device=dev))
            labels, labels_new = labels_new, labels
            batch = 1                                         # a relocating fit is checked every iteration from here on
            if finished:
                break
            continue
        if s_host[0]:
            break
    n_iter = int(state[1].item())
    # labels of the final centres (equal to the last assignment after strict convergence, sklearn's extra E-step otherwise)
    _lib.call('kmpx_kmeans_assign', _lib.ptr(Xc), n, d, K, _lib.ptr(centers), None, _lib.ptr(labels_new), None, st)
    ev[2].record(); ev[2].synchronize()
    info.update(center_ids=center_ids, kpp_ms=ev[0].elapsed_time(ev[1]), lloyd_ms=ev[1].elapsed_time(ev[2]))
    return labels_new, n_iter, mean_dev, info


def _relocate_empty(Xc, labels, centers_old, sums, cnt, rk=None, n_before=0):
    """sklearn's _relocate_empty_clusters_dense (_k_means_common.pyx:167-211): an empty cluster takes the sample
    farthest from its own centre.  Rare; the distances are one fused device expression, with row-partitioned data
    every rank contributes its local top candidates and all ranks apply the same global selection."""
    torch = _lib.require_cuda()
    empty = np.where(cnt == 0)[0]
    c_old = torch.as_tensor(centers_old, device=Xc.device)
    d2 = ((Xc - c_old[labels.long()]) ** 2).sum(dim=1)
    k = min(len(empty), d2.numel())
    top = torch.topk(d2, k)
    rec = torch.zeros((len(empty), Xc.shape[1] + 3), dtype=torch.float64, device=Xc.device)     # [d2, global row, label, x]
    rec[:k, 0] = top.values
    rec[:k, 1] = (top.indices + n_before).double()
    rec[:k, 2] = labels[top.indices].double()
    rec[:k, 3:] = Xc[top.indices]
    if rk is not None and rk.world > 1:
        rec[k:, 0] = -1.0
        rec = rk.gather(rec).reshape(-1, rec.shape[1])
    rec = rec.cpu().numpy()
    order = np.lexsort((rec[:, 1], -rec[:, 0]))[:len(empty)]       # farthest first, ties by global row index
    rec = rec[order]
    if rec[0, 0] == 0.0:
        return sums, cnt
    for e, r in zip(empty, rec):
        old, x = int(r[2]), r[3:]
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
        self.fit_device(torch.as_tensor(data, device='cuda'))

    def fit_device(self, X, group=None, n_total=None, n_before=0, timings=None) -> None:
        """fit() on a CUDA tensor X (n, d); with `group` the rows are partitioned over its ranks (this rank holds rows
        n_before .. n_before + n of n_total): k-means initialisation and EM exchange only packed sums over NCCL, every
        rank ends with identical parameters.  timings: optional dict that receives the device times of init / EM."""
        torch = _lib.require_cuda()
        n, d = X.shape
        K = self.n_components
        n_total = int(n_total if n_total is not None else n)
        if n_total < K:
            raise ValueError(f"Expected n_samples >= n_components but got n_components = {K}, n_samples = {n_total}")
        if not 1 <= d <= 8 or not 1 <= K <= 128:
            raise ValueError(f"the device EM kernels cover 1 <= n_features <= 8 and 1 <= n_components <= 128 "
                             f"(got n_features = {d}, n_components = {K})")
        self.n_features = d
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timings is not None else None
        if ev:
            ev[0].record()
        rs = np.random.RandomState(420)
        labels, km_iter, shift, km_info = kmeans_init_labels(X, K, rs, group=group, n_total=n_total, n_before=n_before)
        if ev:
            ev[1].record()
        eng = EMEngine(X, K, self.diag_reg_factor, shift=shift, group=group, n_total=n_total)
        eng.init_from_labels(labels)
        n_iter, lower, converged = eng.run(tol=1e-3, max_iter=100)
        if ev:
            ev[2].record(); ev[2].synchronize()
            timings.update(init_ms=ev[0].elapsed_time(ev[1]), em_ms=ev[1].elapsed_time(ev[2]), kmeans_iterations=km_iter,
                           em_iterations=n_iter, **{k: v for k, v in km_info.items() if k != 'center_ids'})
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
