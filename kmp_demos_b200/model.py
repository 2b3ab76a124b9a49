"""``KMP`` - drop-in for the reference's ``kmp.model.KMP`` (kmp/model/KMP.py) whose numerical bodies are
calls into libkmpx.so (sm_100a CUDA kernels, include/kmpx.h).

Same constructor, methods, argument shapes and host-visible attributes as the reference:
``KMP(l, alpha, sigma_f, adaptation_tol, time_driven_kernel, verbose)``, ``fit(X (I,N), mu (O,N), var (O,O,N))``,
``predict(s (I,M)) -> (xi (O,M), sigma (O,O,M))``, ``set_waypoint(s, xi, sigma)``; attributes ``s, xi, sigma, N, O,
l, alpha, sigma_f, tol, time_driven_kernel, kl_divergence, _estimator``.

How the path maps to kernels (reference file:line -> entry point):
  KMP.py:146-156  N^2 Python loop building K (x) I + l*Sigma   -> kmpx_kernel_build   (K1)
  KMP.py:157      numpy.linalg.inv (LU)                        -> kmpx_potrf + kmpx_trtri (K2/K3); non-symmetric
                                                                  Sigma blocks (demos.py:251,285) -> kmpx_inverse
  KMP.py:178-190  per-query loop, k@inv@Y, k@inv@k.T           -> kmpx_solve_alpha + kmpx_kmp_predict (K4)
"""
from __future__ import annotations

import copy
import math
import logging
from typing import Tuple

import numpy as np

from . import _lib

logging.basicConfig(
    level=logging.INFO,
    format="%(asctime)s.%(msecs)03d %(levelname)s %(message)s",
    datefmt="%Y-%m-%d,%H:%M:%S",
)


USE_FUSED_FIT = False


def _roundup(v, m):
    return (v + m - 1) // m * m


class _Factorisation:
    """Device state of one fitted system in the component-major order."""

    def __init__(self, F, lda, kind, w, alpha, s_dev, N, I, O):
        self.F, self.lda, self.kind, self.w, self.alpha = F, lda, kind, w, alpha
        self.s_dev, self.N, self.I, self.O = s_dev, N, I, O


def factorise(s, xi, sigma, lam, sigma_f, time_driven, layout=_lib.LAYOUT_COMPONENT, force_general=False):
    """Builds and factorises one KMP system on the device.  s (I,N), xi (O,N), sigma (O,O,N) host arrays.
    Returns a _Factorisation with F = L^{-1} (kind 0) or A^{-1} (kind 1, non-symmetric blocks)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device('cuda')
    I, N = s.shape
    O = xi.shape[0]
    if time_driven and O % 2:
        raise ValueError('time_driven_kernel needs an even number of outputs (KMP.py:125)')
    n = int(lib.kmpx_sys_size(N, O, layout))
    lda = _roundup(n, 8)
    s_dev = torch.as_tensor(np.ascontiguousarray(s.T), dtype=torch.float64, device=dev)                     # (N,I)
    y_dev = torch.as_tensor(np.ascontiguousarray(xi.T), dtype=torch.float64, device=dev)                    # (N,O)
    g_host = np.ascontiguousarray(np.transpose(sigma, (2, 0, 1)))                                            # (N,O,O)
    g_dev = torch.as_tensor(g_host, dtype=torch.float64, device=dev)
    # material asymmetry of the stored covariance blocks => general (LU-like) path, SURVEY.md H2
    scale = float(np.max(np.abs(g_host))) if g_host.size else 0.0
    asym = float(np.max(np.abs(g_host - np.transpose(g_host, (0, 2, 1))))) if g_host.size else 0.0
    general = force_general or (asym > 1e-12 * scale)
    A = torch.empty((n, lda), dtype=torch.float64, device=dev)
    st = _lib.stream()
    # symmetric blocks, plain kernel, O <= 2, N <= ~204 (cfg 1 / cfg 3): build + Cholesky + inverse in ONE shared-memory-
    # resident kernel (factor_struct.cuh) - opt-in through USE_FUSED_FIT, like KMPBatch.fused_fit
    fused = USE_FUSED_FIT and (not general) and (not time_driven) and layout == _lib.LAYOUT_COMPONENT and \
        bool(lib.kmpx_factor_struct_applicable(N, I, O))
    if not fused:
        _lib.call('kmpx_kernel_build', _lib.ptr(s_dev), _lib.ptr(g_dev), None, _lib.ptr(A), 1, N, I, O, lda, n * lda,
                  float(sigma_f), float(lam), int(bool(time_driven)), layout,
                  _lib.UPLO_FULL if general else _lib.UPLO_LOWER, st)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    w = torch.zeros(lda, dtype=torch.float64, device=dev)
    alpha = torch.empty((N, O), dtype=torch.float64, device=dev)
    if not general:
        if fused:
            _lib.call('kmpx_factor_struct', _lib.ptr(s_dev), _lib.ptr(g_dev), None, _lib.ptr(A), lda, n * lda, N, I, O, 1,
                      float(sigma_f), float(lam), _lib.ptr(info), st)
            if int(info.item()) != 0:
                return factorise(s, xi, sigma, lam, sigma_f, time_driven, layout, force_general=True)
        else:
            nbytes = max(int(lib.kmpx_potrf_workspace_bytes(n, 1)), int(lib.kmpx_trtri_workspace_bytes(n, 1)))
            ws = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device=dev)
            _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, st)
            if int(info.item()) != 0:
                # not positive definite: the reference's LU inverse would still go through -> general path
                return factorise(s, xi, sigma, lam, sigma_f, time_driven, layout, force_general=True)
            _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(ws), nbytes, st)
        _lib.call('kmpx_solve_alpha', _lib.ptr(A), lda, n * lda, None, N, O, 1, layout, _lib.ptr(y_dev),
                  _lib.ptr(w), lda, _lib.ptr(alpha), st)
        kind = 0
    else:
        Ainv = None
        if not force_general:
            Ainv = _woodbury_inverse(A, s_dev, g_host, N, I, O, n, lda, lam, sigma_f, time_driven, layout)
        A = Ainv if Ainv is not None else general_inverse(A, n, lda)
        _lib.call('kmpx_apply_inverse', _lib.ptr(A), lda, n * lda, None, N, O, 1, layout, _lib.ptr(y_dev),
                  _lib.ptr(w), lda, _lib.ptr(alpha), st)
        kind = 1
    return _Factorisation(A, lda, kind, w, alpha, s_dev, N, I, O)


def _woodbury_inverse(A, s_dev, g_host, N, I, O, n, lda, lam, sigma_f, time_driven, layout, newton_steps=1):
    """Inverse of a system whose covariance blocks are non-symmetric at a FEW points (the un-listed via-point covariance
    of demos.py:251,285) without the O(n) sequential pivots of Gauss-Jordan: A = S + P W P^T with S the
    system of the symmetrised blocks (Cholesky kernels: potrf, trtri, Linv^T Linv on the DMMA GEMM) and W = lam * skew
    parts on the r = O * (#asymmetric points) rows P selects, so
        A^-1 = S^-1 - S^-1 P W (I + P^T S^-1 P W)^-1 P^T S^-1          (r x r solve on the host),
    followed by Newton-Schulz refinement X <- X + X (I - A X) against the full non-symmetric A, which restores the
    accuracy of a backward-stable LU inverse (what numpy.linalg.inv delivers at KMP.py:157).  Returns None when the
    symmetrised system is not positive definite or too many points are asymmetric (the caller falls back to
    Gauss-Jordan).  A: the full non-symmetric system (n x lda, device); returns A^-1 in the same layout."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = A.device
    st = _lib.stream()
    skew = 0.5 * (g_host - np.transpose(g_host, (0, 2, 1)))
    scale = float(np.max(np.abs(g_host)))
    pts = np.where(np.max(np.abs(skew), axis=(1, 2)) > 1e-12 * scale)[0]
    if len(pts) == 0 or len(pts) * O > 64:
        return None
    g_sym = torch.as_tensor(np.ascontiguousarray(g_host - skew), dtype=torch.float64, device=dev)
    S = torch.empty((n, lda), dtype=torch.float64, device=dev)
    _lib.call('kmpx_kernel_build', _lib.ptr(s_dev), _lib.ptr(g_sym), None, _lib.ptr(S), 1, N, I, O, lda, n * lda,
              float(sigma_f), float(lam), int(bool(time_driven)), layout, _lib.UPLO_LOWER, st)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = max(int(lib.kmpx_potrf_workspace_bytes(n, 1)), int(lib.kmpx_trtri_workspace_bytes(n, 1)))
    ws = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device=dev)
    _lib.call('kmpx_potrf', _lib.ptr(S), lda, n * lda, None, n, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, st)
    if int(info.item()) != 0:
        return None
    _lib.call('kmpx_trtri', _lib.ptr(S), lda, n * lda, None, n, 1, _lib.ptr(ws), nbytes, st)
    S[:, :n].tril_()                                               # small systems keep garbage above the diagonal
    S[:, n:].zero_()
    X = torch.zeros((n, lda), dtype=torch.float64, device=dev)
    _lib.call('kmpx_dgemm', 1, 0, n, n, n, 1.0, _lib.ptr(S), lda, 0, _lib.ptr(S), lda, 0, 0.0, _lib.ptr(X), lda, 0, 1, 0, st)   # S^-1
    # rows of the system the asymmetric points occupy, in the layout's order, and W on them
    Ns = (N + 1) // 2 * 2
    row_of = (lambda i, a: i * O + a) if layout == _lib.LAYOUT_POINT else (lambda i, a: a * Ns + i)
    idx = np.array([row_of(int(i), a) for i in pts for a in range(O)], dtype=np.int64)
    r = len(idx)
    W = np.zeros((r, r))
    for q, i in enumerate(pts):
        W[q * O:(q + 1) * O, q * O:(q + 1) * O] = lam * skew[i]
    idx_d = torch.as_tensor(idx, device=dev)
    Sc = X[:, idx_d][:n].contiguous()                              # S^-1 P   (n x r)
    Sr = X[idx_d, :].contiguous()                                  # P^T S^-1 (r x lda)
    core = Sr[:, idx_d].cpu().numpy()                              # P^T S^-1 P
    M = W @ np.linalg.inv(np.eye(r) + core @ W)
    r2 = r + (r & 1)                                               # even leading dimensions for the GEMM
    Mp = np.zeros((r2, r2)); Mp[:r, :r] = M
    Scp = torch.zeros((n, r2), dtype=torch.float64, device=dev); Scp[:, :r] = Sc
    left = torch.empty((n, r2), dtype=torch.float64, device=dev)
    Md = torch.as_tensor(Mp, device=dev)
    _lib.call('kmpx_dgemm', 0, 0, n, r2, r2, 1.0, _lib.ptr(Scp), r2, 0, _lib.ptr(Md), r2, 0, 0.0, _lib.ptr(left), r2, 0, 1, 0, st)
    right = torch.zeros((r2, lda), dtype=torch.float64, device=dev)
    right[:r] = Sr
    _lib.call('kmpx_dgemm', 0, 0, n, n, r2, -1.0, _lib.ptr(left), r2, 0, _lib.ptr(right), lda, 0, 1.0, _lib.ptr(X), lda, 0, 1, 0, st)
    R = torch.empty_like(X)
    for _ in range(newton_steps):
        R.zero_()
        R[:, :n].fill_diagonal_(1.0)
        _lib.call('kmpx_dgemm', 0, 0, n, n, n, -1.0, _lib.ptr(A), lda, 0, _lib.ptr(X), lda, 0, 1.0, _lib.ptr(R), lda, 0, 1, 0, st)
        X0 = X.clone()
        _lib.call('kmpx_dgemm', 0, 0, n, n, n, 1.0, _lib.ptr(X0), lda, 0, _lib.ptr(R), lda, 0, 1.0, _lib.ptr(X), lda, 0, 1, 0, st)
    return X


def general_inverse(A, n, lda, newton_steps=2):
    """In-place inverse of a general (non-symmetric) n x n device matrix: Gauss-Jordan with partial pivoting
    (kmpx_inverse) followed by Newton-Schulz refinement X <- X + X (I - A X) on the DMMA GEMM, which brings the
    result to the accuracy of a backward-stable LU inverse (what numpy.linalg.inv delivers at KMP.py:157)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = A.device
    st = _lib.stream()
    A0 = A.clone()
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = int(lib.kmpx_inverse_workspace_bytes(n, 1))
    ws = torch.empty(max(nbytes, 8) // 8 + 1, dtype=torch.float64, device=dev)
    _lib.call('kmpx_inverse', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, st)
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError('Singular matrix')      # what numpy.linalg.inv raises
    R = torch.empty_like(A)
    for _ in range(newton_steps):
        R.zero_()
        R[:, :n].fill_diagonal_(1.0)
        _lib.call('kmpx_dgemm', 0, 0, n, n, n, -1.0, _lib.ptr(A0), lda, 0, _lib.ptr(A), lda, 0, 1.0, _lib.ptr(R), lda, 0, 1, 0, st)
        X = A.clone()
        _lib.call('kmpx_dgemm', 0, 0, n, n, n, 1.0, _lib.ptr(X), lda, 0, _lib.ptr(R), lda, 0, 1.0, _lib.ptr(A), lda, 0, 1, 0, st)
    return A


class KMP:
    """Trajectory imitation and adaptation using Kernelized Movement Primitives (B200 implementation).

    Parameters (identical to the reference, KMP.py:39-47)
    ----------
    l : float, default=0.5            lambda regularisation factor.
    alpha : float, default=40         coefficient of the covariance prediction.
    sigma_f : float, default=1        kernel coefficient.
    adaptation_tol : float, default=0.0005   distance below which a via-point replaces a stored point.
    time_driven_kernel : bool, default=True  finite-difference position/velocity kernel.
    verbose : bool, default=False
    """

    def __init__(self, l: float = 0.5, alpha: float = 40, sigma_f: float = 1.0, adaptation_tol: float = 0.0005,
                 time_driven_kernel: bool = True, verbose: bool = False) -> None:
        self.l = l
        self.alpha = alpha
        self.sigma_f = sigma_f
        self.kl_divergence = None
        self.tol = adaptation_tol
        self.time_driven_kernel = time_driven_kernel
        self._logger = logging.getLogger(__name__)
        self._logger.setLevel(level=logging.DEBUG if verbose else logging.INFO)
        self._fact = None
        self._estimator_cache = None

    # ------------------------------------------------------------------ fit (KMP.py:130-158)
    def fit(self, X: np.ndarray, mu: np.ndarray, var: np.ndarray) -> None:
        """Builds K (x) I + l*Sigma on the device and factorises it.  X (I,N), mu (O,N), var (O,O,N);
        the inputs are deep-copied like the reference does (KMP.py:142-144)."""
        self.s = copy.deepcopy(X)
        self.xi = copy.deepcopy(mu)
        self.sigma = copy.deepcopy(var)
        self.O, self.N = self.xi.shape
        self._fact = factorise(np.asarray(self.s, dtype=np.float64), np.asarray(self.xi, dtype=np.float64),
                               np.asarray(self.sigma, dtype=np.float64), self.l, self.sigma_f,
                               self.time_driven_kernel)
        self._estimator_cache = None
        self._logger.info("KMP fit done.")

    # ------------------------------------------------------------------ predict (KMP.py:160-194)
    def predict(self, s: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Mean (O,M) and covariance (O,O,M) for the inputs s (I,M)."""
        torch = _lib.require_cuda()
        if self._fact is None:
            raise AttributeError("KMP.predict called before fit")
        s = np.asarray(s, dtype=np.float64)
        sq = torch.as_tensor(np.ascontiguousarray(s.T), dtype=torch.float64, device=torch.device('cuda'))
        xi, sg = self.predict_device(sq)
        self._logger.info("KMP predict done.")
        return xi.cpu().numpy(), sg.cpu().numpy()

    def predict_device(self, sq, xi=None, sg=None):
        """predict() on device tensors: sq (M,I) float64 CUDA -> xi (O,M), sg (O,O,M) CUDA tensors (optionally
        caller-provided).  No host synchronisation."""
        torch = _lib.require_cuda()
        f = self._fact
        M = sq.shape[0]
        dev = sq.device
        if xi is None:
            xi = torch.empty((f.O, M), dtype=torch.float64, device=dev)
        if sg is None:
            sg = torch.empty((f.O, f.O, M), dtype=torch.float64, device=dev)
        n = f.F.shape[0]
        if f.kind == 0 and not self.time_driven_kernel and f.O <= 2 and n > _lib.SMALL_MAX:
            # one large system: kappa + DMMA GEMMs + column reductions (predict_gemm.cu), queries in chunks of <= 8192
            lib = _lib.load()
            mc = min((M + 127) // 128 * 128, 8192)
            nbytes = int(lib.kmpx_kmp_predict_gemm_workspace_bytes(f.N, f.O, mc))
            if getattr(self, '_gemm_ws', None) is None or self._gemm_ws.numel() * 8 < nbytes or self._gemm_ws.device != dev:
                self._gemm_ws = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
            _lib.call('kmpx_kmp_predict_gemm', _lib.ptr(f.F), f.lda, _lib.ptr(f.w), _lib.ptr(f.s_dev), f.N, f.I, f.O,
                      _lib.ptr(sq), M, float(self.sigma_f), float(self.alpha), _lib.ptr(self._gemm_ws), nbytes, _lib.ptr(xi),
                      _lib.ptr(sg), _lib.stream())
        else:
            _lib.call('kmpx_kmp_predict', _lib.ptr(f.F), f.lda, n * f.lda, f.kind, _lib.ptr(f.w), f.lda,
                      _lib.ptr(f.s_dev), None, f.N, f.I, f.O, 1, _lib.ptr(sq), 0, M, float(self.sigma_f),
                      float(self.alpha), int(bool(self.time_driven_kernel)), _lib.ptr(xi), _lib.ptr(sg), _lib.stream())
        return xi, sg

    # ------------------------------------------------------------------ one large system, queries sharded over the GPUs
    def broadcast_fit(self, src: int = 0, group=None) -> None:
        """Ships the fitted state (L^-1 or A^-1, w, alpha, inputs) of rank `src` to every rank of the process group
        over NCCL, so that only one GPU pays for the factorisation (1.07 GB at n = 16384).  Every rank must have
        called fit() on a problem of the same shape, or pass through here with an un-fitted model of rank != src
        after set_shape_like(); the simplest use is: every rank builds the model, rank `src` fits, all call this."""
        torch = _lib.require_cuda()
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        rank = dist.get_rank(group)
        meta = [None]
        if rank == src:
            f = self._fact
            meta = [dict(shape=tuple(f.F.shape), lda=f.lda, kind=f.kind, N=f.N, I=f.I, O=f.O, w=tuple(f.w.shape),
                         alpha=tuple(f.alpha.shape), s=self.s, xi=self.xi, sigma=self.sigma)]
        dist.broadcast_object_list(meta, src=src, group=group)
        m = meta[0]
        dev = torch.device('cuda', torch.cuda.current_device())
        if rank != src:
            e = lambda shp: torch.empty(shp, dtype=torch.float64, device=dev)  # noqa: E731
            self.s, self.xi, self.sigma = m['s'], m['xi'], m['sigma']
            self.O, self.N = self.xi.shape
            self._fact = _Factorisation(e(m['shape']), m['lda'], m['kind'], e(m['w']), e(m['alpha']), e((m['N'], m['I'])),
                                        m['N'], m['I'], m['O'])
            self._estimator_cache = None
        f = self._fact
        for t in (f.F, f.w, f.alpha, f.s_dev):
            dist.broadcast(t, src=src, group=group)

    def predict_sharded(self, s: np.ndarray, group=None) -> Tuple[np.ndarray, np.ndarray]:
        """predict() with the queries split contiguously over the ranks of the process group (each rank needs the
        fitted state: fit() everywhere, or fit() on one rank + broadcast_fit()); the shards are all-gathered on the
        device (NCCL), so every rank returns the full (O,M) mean and (O,O,M) covariance after ONE device-to-host copy.
        No collective touches the arithmetic."""
        torch = _lib.require_cuda()
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.predict(s)
        s = np.asarray(s, dtype=np.float64)
        dev = torch.device('cuda', torch.cuda.current_device())
        sq = torch.as_tensor(np.ascontiguousarray(s.T), dtype=torch.float64, device=dev)
        out = self.predict_sharded_device(sq, group)
        O = self.O
        h = out.cpu().numpy()
        return h[:O].copy(), h[O:].reshape(O, O, -1).copy()

    def predict_sharded_device(self, sq, group=None):
        """Device form of predict_sharded: sq (M,I) CUDA tensor (the same on every rank) -> (O + O*O, M) CUDA tensor
        [means; covariances] holding all queries on every rank."""
        torch = _lib.require_cuda()
        import torch.distributed as dist
        from .dist import shard_range
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        M, O = sq.shape[0], self.O
        lo, hi = shard_range(M, rank, world)
        per = shard_range(M, 0, world)[1]                       # rank 0 holds the longest shard
        dev = sq.device
        gath = torch.empty((world, O + O * O, per), dtype=torch.float64, device=dev)
        mine = gath[rank]
        if hi > lo:
            # the kernels write straight into this rank's slot of the gather buffer (rows of `per` columns)
            xi_l, sg_l = self.predict_device(sq[lo:hi].contiguous())
            mine[:O, :hi - lo] = xi_l
            mine[O:, :hi - lo] = sg_l.reshape(O * O, -1)
        dist.all_gather_into_tensor(gath.view(-1), mine.reshape(-1).clone(), group=group)
        if M == per * world:
            return gath.permute(1, 0, 2).reshape(O + O * O, M)
        out = torch.empty((O + O * O, M), dtype=torch.float64, device=dev)
        for r in range(world):
            a, b_ = shard_range(M, r, world)
            out[:, a:b_] = gath[r, :, :b_ - a]
        return out

    def predict_mean(self, s: np.ndarray) -> np.ndarray:
        """Mean only, from alpha = (K + l*Sigma)^-1 mu: O(N*M) kernel evaluations instead of the GEMM."""
        torch = _lib.require_cuda()
        f = self._fact
        s = np.asarray(s, dtype=np.float64)
        M = s.shape[1]
        dev = torch.device('cuda')
        sq = torch.as_tensor(np.ascontiguousarray(s.T), dtype=torch.float64, device=dev)
        xi = torch.empty((f.O, M), dtype=torch.float64, device=dev)
        _lib.call('kmpx_kmp_predict_mean', _lib.ptr(f.alpha), _lib.ptr(f.s_dev), None, f.N, f.I, f.O, 1, _lib.ptr(sq), 0, M,
                  float(self.sigma_f), int(bool(self.time_driven_kernel)), _lib.ptr(xi), _lib.stream())
        return xi.cpu().numpy()

    # ------------------------------------------------------------------ set_waypoint (KMP.py:57-91)
    def set_waypoint(self, s, xi, sigma) -> None:
        """Inserts via-points: the nearest stored input (L2, first minimum, only among the ``self.N`` points of
        the last fit) is overwritten when closer than ``tol`` (strict), otherwise the point is appended; then
        the model is refitted.  ``s[j]``, ``xi[j]`` and ``sigma[j]`` are used exactly as indexed by the caller,
        so the reference's un-listed forms (demos.py:251,285) broadcast the same way."""
        for j in range(len(s)):
            d = np.linalg.norm(self.s[:, :self.N] - np.asarray(s[j], dtype=float).reshape(-1, 1), axis=0) \
                if self.s.shape[0] > 1 else np.abs(self.s[0, :self.N] - s[j])
            d = np.where(np.isnan(d), np.inf, d)          # the reference's `dist < min_dist` scan skips NaN (KMP.py:75-79)
            idx = int(np.argmin(d)) if self.N > 0 else -1
            if idx >= 0 and d[idx] < self.tol:
                self.s[:, idx] = s[j]
                self.xi[:, idx] = xi[j]
                self.sigma[:, :, idx] = sigma[j]
            else:
                self.s = np.concatenate((self.s, np.array(s[j]).reshape(1, -1)), axis=1)
                self.xi = np.concatenate((self.xi, xi[j].T), axis=1)
                self.sigma = np.concatenate((self.sigma, np.expand_dims(sigma[j], 2)), axis=2)
        self.fit(self.s, self.xi, self.sigma)

    # ------------------------------------------------------------------ _estimator (KMP.py:157), lazily
    @property
    def _estimator(self) -> np.ndarray:
        """inv(K (x) I + l*Sigma) in the reference's point-major order, materialised on first access
        (the prediction path never needs it)."""
        if self._estimator_cache is None:
            self._estimator_cache = explicit_inverse(np.asarray(self.s, float), np.asarray(self.sigma, float), self.l,
                                                     self.sigma_f, self.time_driven_kernel)
        return self._estimator_cache

    def system_matrix(self, layout=_lib.LAYOUT_POINT) -> np.ndarray:
        """K (x) I + l*Sigma as built by kmpx_kernel_build (point-major = the reference's ``k``, KMP.py:146-156)."""
        return build_matrix(np.asarray(self.s, float), np.asarray(self.sigma, float), self.l, self.sigma_f,
                            self.time_driven_kernel, layout)

    def KL_divergence(self, xi, sigma, xi_ref, sigma_ref) -> float:
        """The reference's aggregate of point-wise "KL" terms (KMP.py:196-212; its only call site, KMP.py:192, is
        commented out upstream): for each of the self.N stored points, with N(mu, S) the normal density,
        kl_i = q_i log(q_i / p_i), p_i = N(xi_i; xi_i, sigma_i), q_i = N(xi_i; xi_ref_i, sigma_ref_i); the vector is
        normalised by its 2-norm and averaged.  Host arithmetic (O x O blocks), densities as scipy's
        multivariate_normal evaluates them (eigendecomposition, pseudo-determinant)."""
        def pdf(x, mu, cov):
            w, v = np.linalg.eigh(np.atleast_2d(cov))
            eps = 1e6 * np.finfo(float).eps * np.max(np.abs(w))
            if np.min(w) < -eps:
                raise ValueError('the input matrix must be positive semidefinite')      # scipy's message
            keep = w > eps
            d = w[keep]
            dev = np.atleast_1d(x) - np.atleast_1d(mu)
            maha = np.sum(((dev @ v[:, keep]) / np.sqrt(d)) ** 2)
            return math.exp(-0.5 * (int(keep.sum()) * math.log(2.0 * math.pi) + np.sum(np.log(d)) + maha))
        kl = np.empty(self.N)
        for i in range(self.N):
            p_i = pdf(xi[:, i], xi[:, i], sigma[:, :, i])
            q_i = pdf(xi[:, i], xi_ref[:, i], sigma_ref[:, :, i])
            kl[i] = q_i * np.log(q_i / p_i)
        kl /= np.linalg.norm(kl)
        return float(np.mean(kl))


def build_matrix(s, sigma, lam, sigma_f, time_driven, layout=_lib.LAYOUT_POINT):
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device('cuda')
    I, N = s.shape
    O = sigma.shape[0]
    n = int(lib.kmpx_sys_size(N, O, layout))
    lda = _roundup(n, 8)
    s_dev = torch.as_tensor(np.ascontiguousarray(s.T), dtype=torch.float64, device=dev)
    g_dev = torch.as_tensor(np.ascontiguousarray(np.transpose(sigma, (2, 0, 1))), dtype=torch.float64, device=dev)
    A = torch.empty((n, lda), dtype=torch.float64, device=dev)
    _lib.call('kmpx_kernel_build', _lib.ptr(s_dev), _lib.ptr(g_dev), None, _lib.ptr(A), 1, N, I, O, lda, n * lda,
              float(sigma_f), float(lam), int(bool(time_driven)), layout, _lib.UPLO_FULL, _lib.stream())
    return A[:, :n].cpu().numpy()


def explicit_inverse(s, sigma, lam, sigma_f, time_driven):
    """Point-major inverse on the device: potrf + trtri + Linv^T Linv (DMMA GEMM), or Gauss-Jordan when the
    covariance blocks are not symmetric."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device('cuda')
    I, N = s.shape
    O = sigma.shape[0]
    n = N * O
    lda = _roundup(n, 8)
    s_dev = torch.as_tensor(np.ascontiguousarray(s.T), dtype=torch.float64, device=dev)
    g_host = np.ascontiguousarray(np.transpose(sigma, (2, 0, 1)))
    g_dev = torch.as_tensor(g_host, dtype=torch.float64, device=dev)
    A = torch.empty((n, lda), dtype=torch.float64, device=dev)
    st = _lib.stream()
    _lib.call('kmpx_kernel_build', _lib.ptr(s_dev), _lib.ptr(g_dev), None, _lib.ptr(A), 1, N, I, O, lda, n * lda,
              float(sigma_f), float(lam), int(bool(time_driven)), _lib.LAYOUT_POINT, _lib.UPLO_FULL, st)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    scale = float(np.max(np.abs(g_host)))
    asym = float(np.max(np.abs(g_host - np.transpose(g_host, (0, 2, 1)))))
    if asym <= 1e-12 * scale:
        nbytes = max(int(lib.kmpx_potrf_workspace_bytes(n, 1)), int(lib.kmpx_trtri_workspace_bytes(n, 1)))
        ws = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device=dev)
        A0 = A.clone()
        _lib.call('kmpx_potrf', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(info), _lib.ptr(ws), nbytes, st)
        if int(info.item()) == 0:
            _lib.call('kmpx_trtri', _lib.ptr(A), lda, n * lda, None, n, 1, _lib.ptr(ws), nbytes, st)
            Linv = torch.tril(A[:, :n]).contiguous()
            ldl = Linv.stride(0)
            if ldl % 2:
                Linv = torch.nn.functional.pad(Linv, (0, 1)).contiguous()
                ldl = Linv.stride(0)
            out = torch.empty((n, lda), dtype=torch.float64, device=dev)
            _lib.call('kmpx_dgemm', 1, 0, n, n, n, 1.0, _lib.ptr(Linv), ldl, 0, _lib.ptr(Linv), ldl, 0, 0.0,
                      _lib.ptr(out), lda, 0, 1, 0, st)
            return out[:, :n].cpu().numpy()
        A = A0
    A = general_inverse(A, n, lda)
    return A[:, :n].cpu().numpy()
