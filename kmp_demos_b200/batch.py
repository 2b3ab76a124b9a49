"""Batched KMP adaptation: many independent (base trajectory, via-point set) problems through one
set_waypoint -> fit -> predict pipeline on the device (SURVEY.md 8d cfg 3).

One problem = what the reference does with
    kmp.fit(t, mu, sigma); kmp.set_waypoint(via_t, via_p, via_var); kmp.predict(queries)
(KMP.py:57-91,130-194); here B of them run per kernel launch:
    kmpx_waypoint_insert -> kmpx_kernel_build -> kmpx_potrf -> kmpx_trtri -> kmpx_solve_alpha -> kmpx_kmp_predict
Problems are independent, so a multi-GPU run just gives every rank a contiguous slice (no collective).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _roundup(v, m):
    return (v + m - 1) // m * m


WARP_MIN_BATCH = 4 * 148     # problems per call from which the warp-per-problem factor kernels are taken ('auto')


class KMPBatch:
    """Batched fit+predict of KMP problems that share base trajectories and a query grid.

    Parameters mirror ``KMP`` (l, alpha, sigma_f, adaptation_tol, time_driven_kernel).  ``chunk`` bounds the
    number of systems resident at once (each is n^2 doubles)."""

    def __init__(self, l=0.5, alpha=40.0, sigma_f=1.0, adaptation_tol=0.0005, time_driven_kernel=False, chunk=5328):
        self.l, self.alpha, self.sigma_f = float(l), float(alpha), float(sigma_f)
        self.tol, self.time_driven_kernel, self.chunk = float(adaptation_tol), bool(time_driven_kernel), int(chunk)
        self._bufs = {}
        self._base = None
        # plain kernel, O <= 2, N <= ~204: one fused shared-memory-resident fit kernel (factor_struct.cuh) instead of
        # build + potrf + trtri.  Correct and tested, but measured ~20 % SLOWER than the TMA pipeline on the cfg-3 shape
        # (DESIGN.md 3.4), hence opt-in
        self.fused_fit = False
        # Cholesky + triangular inverse of the batched path: 'cta' = one CTA per problem (factor_tma.cuh), 'warp' = one
        # warp per problem, twelve independent problems per SM (factor_wpp.cuh; 1.8x faster on a cfg-3 chunk, but it needs
        # a few problems per warp slot to pay), 'auto' = 'warp' when the call holds >= WARP_MIN_BATCH problems.  The
        # choice is made once per call from the total batch, never per chunk, so that a problem takes the same kernel -
        # and yields the same bits - whichever chunk it falls into.
        self.factor_mode = 'auto'
        # fused mean + covariance predict: 'tile' = one CTA per tile of 32 or 64 queries (predict_tab.cu, the default);
        # 'persistent' = persistent CTAs over tiles of 40 queries with double-buffered kappa panels (predict_pers.cu):
        # built to hide the panel evaluation and the epilogue, correct, but measured 6 % SLOWER on the cfg-3 shape
        #  - opt-in, kept as the measurement behind that paragraph.
        self.predict_mode = 'tile'
        self.events = None        # set to a list to collect (kernel name, start, end) CUDA events per launch

    # ------------------------------------------------------------------ data staging
    def set_base(self, s_base, xi_base, sigma_base):
        """Base trajectories: s_base (L,I,N0), xi_base (L,O,N0), sigma_base (L,O,O,N0) host arrays in the
        reference's layouts (sample index last); copied to the device once."""
        torch = _lib.require_cuda()
        dev = torch.device('cuda')
        s_base = np.asarray(s_base, dtype=np.float64)
        xi_base = np.asarray(xi_base, dtype=np.float64)
        sigma_base = np.asarray(sigma_base, dtype=np.float64)
        L, I, N0 = s_base.shape
        O = xi_base.shape[1]
        self._base = dict(
            L=L, I=I, O=O, N0=N0,
            s=torch.as_tensor(np.ascontiguousarray(np.transpose(s_base, (0, 2, 1))), device=dev),          # (L,N0,I)
            y=torch.as_tensor(np.ascontiguousarray(np.transpose(xi_base, (0, 2, 1))), device=dev),         # (L,N0,O)
            g=torch.as_tensor(np.ascontiguousarray(np.transpose(sigma_base, (0, 3, 1, 2))), device=dev),   # (L,N0,O,O)
        )

    def _launch(self, name, *args):
        if self.events is None:
            _lib.call(name, *args)
            return
        torch = _lib.require_cuda()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call(name, *args)
        e1.record()
        self.events.append((name, e0, e1))

    def _buf(self, name, shape, dtype):
        torch = _lib.require_cuda()
        key = (name, tuple(shape), dtype)
        if key not in self._bufs:
            self._bufs[key] = torch.empty(shape, dtype=dtype, device='cuda')
        return self._bufs[key]

    # ------------------------------------------------------------------ device pipeline
    def run_device(self, base_of, via_s, via_xi, via_sigma, queries, want_cov=True, out_xi=None, out_sigma=None,
                   host_xi=None, host_sigma=None):
        """All arguments are device tensors: base_of int32 (B,), via_s (B,P,I), via_xi (B,P,O), via_sigma (B,P,O,O),
        queries (M,I).  Returns (xi (B,O,M), sigma (B,O,O,M) or None, info (B,) int32, n_pts (B,) int32).
        host_xi / host_sigma: optional pinned host tensors of the output shapes; the results of every chunk are copied
        out on a side stream while the next chunk is computed, and the current stream waits for the last copy."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        b = self._base
        if b is None:
            raise RuntimeError('KMPBatch.set_base must be called first')
        I, O, N0 = b['I'], b['O'], b['N0']
        B, P = via_s.shape[0], via_s.shape[1]
        M = queries.shape[0]
        n_max_pts = N0 + P
        n_max = int(lib.kmpx_sys_size(n_max_pts, O, _lib.LAYOUT_COMPONENT))
        if n_max > _lib.SMALL_MAX:
            raise ValueError(f'batched path handles systems up to {_lib.SMALL_MAX}, got {n_max}')
        lda = _roundup(n_max, 8)
        f64, i32 = torch.float64, torch.int32
        xi = out_xi if out_xi is not None else torch.empty((B, O, M), dtype=f64, device='cuda')
        sg = None
        if want_cov:
            sg = out_sigma if out_sigma is not None else torch.empty((B, O, O, M), dtype=f64, device='cuda')
        # results handed to the caller are allocated per call (never a cached buffer a later call would overwrite)
        info = torch.empty((B,), dtype=i32, device='cuda')
        n_pts = torch.empty((B,), dtype=i32, device='cuda')
        Bc = min(self.chunk, B)
        A = self._buf('A', (Bc, n_max, lda), f64)
        s = self._buf('s', (Bc, n_max_pts, I), f64)
        y = self._buf('y', (Bc, n_max_pts, O), f64)
        g = self._buf('g', (Bc, n_max_pts, O, O), f64)
        w = self._buf('w', (Bc, lda), f64)
        al = self._buf('alpha', (Bc, n_max_pts, O), f64)
        n_sys = self._buf('n_sys', (Bc,), i32)
        st = _lib.stream()
        td = int(self.time_driven_kernel)
        strideA = n_max * lda
        # the covariance blocks handed to this path are symmetric by contract (GMR output, via-point covariances)
        fused_fit = (not self.time_driven_kernel) and self.fused_fit and bool(lib.kmpx_factor_struct_applicable(n_max_pts, I, O))
        if self.factor_mode not in ('auto', 'cta', 'warp'):
            raise ValueError("factor_mode must be 'auto', 'cta' or 'warp'")
        warp_fit = (not fused_fit) and self.factor_mode != 'cta' and (self.factor_mode == 'warp' or B >= WARP_MIN_BATCH) \
            and bool(lib.kmpx_factor_wpp_applicable(_lib.ptr(A), lda, strideA, n_max, Bc))
        struct_o = 2 if (O == 2 and not self.time_driven_kernel) else 0      # [[A00, D], [D, A11]] with D diagonal
        if self.predict_mode not in ('tile', 'persistent'):
            raise ValueError("predict_mode must be 'tile' or 'persistent'")
        predict_entry = 'kmpx_kmp_predict_persistent' if self.predict_mode == 'persistent' else 'kmpx_kmp_predict'
        copy_stream = None
        if host_xi is not None or host_sigma is not None:
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream()
            copy_stream = self._copy_stream
            copy_stream.wait_stream(torch.cuda.current_stream())     # the previous contents of the host buffers are dead
        for c0 in range(0, B, Bc):
            bc = min(Bc, B - c0)
            npc = n_pts[c0:c0 + bc]
            self._launch('kmpx_waypoint_insert', _lib.ptr(b['s']), _lib.ptr(b['y']), _lib.ptr(b['g']), None, N0,
                      _lib.ptr(base_of[c0:]), _lib.ptr(via_s[c0:]), _lib.ptr(via_xi[c0:]), _lib.ptr(via_sigma[c0:]), P,
                      self.tol, _lib.ptr(s), _lib.ptr(y), _lib.ptr(g), _lib.ptr(npc), n_max_pts, I, O, bc, st)
            if fused_fit:
                # build + Cholesky + triangular inverse in one shared-memory-resident kernel (factor_struct.cuh)
                self._launch('kmpx_factor_struct', _lib.ptr(s), _lib.ptr(g), _lib.ptr(npc), _lib.ptr(A), lda, strideA, n_max_pts, I, O,
                             bc, self.sigma_f, self.l, _lib.ptr(info[c0:]), st)
            else:
                self._launch('kmpx_sys_sizes', _lib.ptr(npc), O, _lib.LAYOUT_COMPONENT, bc, _lib.ptr(n_sys), st)
                self._launch('kmpx_kernel_build', _lib.ptr(s), _lib.ptr(g), _lib.ptr(npc), _lib.ptr(A), bc, n_max_pts, I, O, lda,
                             strideA, self.sigma_f, self.l, td, _lib.LAYOUT_COMPONENT, _lib.UPLO_LOWER, st)
                if warp_fit:
                    self._launch('kmpx_potrf_wpp', _lib.ptr(A), lda, strideA, _lib.ptr(n_sys), n_max, bc, _lib.ptr(info[c0:]), struct_o, st)
                    if want_cov:       # w = Linv y' falls out of the inverse (no pass of its own)
                        self._launch('kmpx_trtri_wpp_solve', _lib.ptr(A), lda, strideA, _lib.ptr(n_sys), n_max, bc, struct_o,
                                     _lib.ptr(y), _lib.ptr(npc), n_max_pts, O, _lib.ptr(w), lda, st)
                    else:
                        self._launch('kmpx_trtri_wpp', _lib.ptr(A), lda, strideA, _lib.ptr(n_sys), n_max, bc, struct_o, st)
                else:
                    self._launch('kmpx_potrf', _lib.ptr(A), lda, strideA, _lib.ptr(n_sys), n_max, bc, _lib.ptr(info[c0:]), None, 0, st)
                    self._launch('kmpx_trtri', _lib.ptr(A), lda, strideA, _lib.ptr(n_sys), n_max, bc, None, 0, st)
            # alpha = A^-1 mu is only needed by the mean-only predict; the fused mean+covariance kernel takes w
            if not (warp_fit and want_cov):
                self._launch('kmpx_solve_alpha', _lib.ptr(A), lda, strideA, _lib.ptr(npc), n_max_pts, O, bc,
                      _lib.LAYOUT_COMPONENT, _lib.ptr(y), _lib.ptr(w), lda, None if want_cov else _lib.ptr(al), st)
            # with host outputs the predict of the LAST chunk is launched in four parts, so that only the copy of the last
            # quarter is left when the compute stream is done (the copies of earlier chunks overlap with the next chunk)
            parts = 4 if (copy_stream is not None and c0 + bc >= B and bc >= 64) else 1
            for q0, q1 in ((bc * i // parts, bc * (i + 1) // parts) for i in range(parts)):
                nb = q1 - q0
                if want_cov:
                    self._launch(predict_entry, _lib.ptr(A[q0:]), lda, strideA, 0, _lib.ptr(w[q0:]), lda, _lib.ptr(s[q0:]),
                              _lib.ptr(npc[q0:]), n_max_pts, I, O, nb, _lib.ptr(queries), 0, M, self.sigma_f, self.alpha, td,
                              _lib.ptr(xi[c0 + q0:]), _lib.ptr(sg[c0 + q0:]), st)
                else:
                    self._launch('kmpx_kmp_predict_mean', _lib.ptr(al[q0:]), _lib.ptr(s[q0:]), _lib.ptr(npc[q0:]), n_max_pts, I, O, nb,
                              _lib.ptr(queries), 0, M, self.sigma_f, td, _lib.ptr(xi[c0 + q0:]), st)
                if copy_stream is not None:
                    done = torch.cuda.Event()
                    done.record()
                    copy_stream.wait_event(done)
                    with torch.cuda.stream(copy_stream):
                        if host_xi is not None:
                            host_xi[c0 + q0:c0 + q1].copy_(xi[c0 + q0:c0 + q1], non_blocking=True)
                        if host_sigma is not None and sg is not None:
                            host_sigma[c0 + q0:c0 + q1].copy_(sg[c0 + q0:c0 + q1], non_blocking=True)
        if copy_stream is not None:
            torch.cuda.current_stream().wait_stream(copy_stream)
        return xi, sg, info, n_pts

    # ------------------------------------------------------------------ incremental adaptation
    def run_incremental(self, base_of, via_s, via_xi, via_sigma, queries, want_cov=True, out_xi=None, out_sigma=None,
                        host_xi=None, host_sigma=None):
        """Same arguments and results as run_device, computed by an exact low-rank correction of the per-base
        solution instead of one factorisation per problem (csrc/incremental.cu): the bases are factorised, inverted
        and predicted once per call (direct kernels, batch = number of bases), every problem then costs O(n^2 + n M).
        Plain kernel, O <= 2, at most two via points per problem.  Returns (xi, sigma, info of the BASE systems);
        the caller checks `info` (check_base_info) - the call itself never synchronises with the host."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        b = self._base
        if b is None:
            raise RuntimeError('KMPBatch.set_base must be called first')
        if self.time_driven_kernel:
            raise ValueError('incremental adaptation is implemented for the plain kernel')
        I, O, N0, L = b['I'], b['O'], b['N0'], b['L']
        B, P = via_s.shape[0], via_s.shape[1]
        M = queries.shape[0]
        f64, i32 = torch.float64, torch.int32
        n0 = int(lib.kmpx_sys_size(N0, O, _lib.LAYOUT_COMPONENT))
        lda = _roundup(n0, 8)
        st = _lib.stream()
        # ---- per base: A_b, L^-1, w, base predictions, G = L^-T L^-1, alpha = G y, kappa
        A = self._buf('incA', (L, n0, lda), f64)
        A.zero_()                                   # K1 writes the lower triangle only: L^-1 must carry zeros above it
        G = self._buf('incG', (L, n0, lda), f64)
        n_pts = self._buf('inc_npts', (L,), i32); n_pts.fill_(N0)
        n_sys = self._buf('inc_nsys', (L,), i32)
        info = self._buf('inc_info', (L,), i32)
        w = self._buf('inc_w', (L, lda), f64)
        xi_b = self._buf('inc_xib', (L, O, M), f64)
        sg_b = self._buf('inc_sgb', (L, O, O, M), f64)
        al = self._buf('inc_alpha', (L, n0), f64)
        kap = self._buf('inc_kappa', (L, N0, M), f64)
        strideA = n0 * lda
        self._launch('kmpx_sys_sizes', _lib.ptr(n_pts), O, _lib.LAYOUT_COMPONENT, L, _lib.ptr(n_sys), st)
        self._launch('kmpx_kernel_build', _lib.ptr(b['s']), _lib.ptr(b['g']), _lib.ptr(n_pts), _lib.ptr(A), L, N0, I, O, lda,
                     strideA, self.sigma_f, self.l, 0, _lib.LAYOUT_COMPONENT, _lib.UPLO_LOWER, st)
        self._launch('kmpx_potrf', _lib.ptr(A), lda, strideA, _lib.ptr(n_sys), n0, L, _lib.ptr(info), None, 0, st)
        self._launch('kmpx_trtri', _lib.ptr(A), lda, strideA, _lib.ptr(n_sys), n0, L, None, 0, st)
        self._launch('kmpx_solve_alpha', _lib.ptr(A), lda, strideA, _lib.ptr(n_pts), N0, O, L, _lib.LAYOUT_COMPONENT,
                     _lib.ptr(b['y']), _lib.ptr(w), lda, None, st)
        self._launch('kmpx_kmp_predict', _lib.ptr(A), lda, strideA, 0, _lib.ptr(w), lda, _lib.ptr(b['s']), _lib.ptr(n_pts),
                     N0, I, O, L, _lib.ptr(queries), 0, M, self.sigma_f, self.alpha, 0, _lib.ptr(xi_b), _lib.ptr(sg_b), st)
        self._launch('kmpx_dgemm', 1, 0, n0, n0, n0, 1.0, _lib.ptr(A), lda, strideA, _lib.ptr(A), lda, strideA, 0.0,
                     _lib.ptr(G), lda, strideA, L, 0, st)
        self._launch('kmpx_inc_base', _lib.ptr(G), lda, strideA, _lib.ptr(b['y']), _lib.ptr(b['s']), N0, I, O, L,
                     _lib.ptr(queries), M, self.sigma_f, _lib.ptr(al), _lib.ptr(kap), st)
        # ---- per problem
        xi = out_xi if out_xi is not None else torch.empty((B, O, M), dtype=f64, device='cuda')
        sg = None
        if want_cov:
            sg = out_sigma if out_sigma is not None else torch.empty((B, O, O, M), dtype=f64, device='cuda')
        # small chunks when the results stream out to the host: the step is bound by that copy, and only the first chunk's
        # compute and the last chunk's copy are not overlapped
        Bc = min(self.chunk if (host_xi is None and host_sigma is None) else min(self.chunk, 1184), B)
        # problems sorted by base (the usual case): the appended columns of all problems of a base come from one GEMM.
        # The grouping is a launch-structure parameter (HOST arrays in the C ABI); it is read back once per distinct
        # base_of tensor, not once per call
        gkey = (base_of.data_ptr(), base_of._version, B)
        if getattr(self, '_group_key', None) != gkey:
            self._group_key, self._group_host = gkey, base_of.cpu().numpy()
        base_host = self._group_host
        grouped = bool(np.all(base_host[1:] >= base_host[:-1]))
        nws = int((lib.kmpx_inc_grouped_workspace_doubles if grouped else lib.kmpx_inc_workspace_doubles)(N0, O, Bc))
        ws = self._buf('inc_ws', (nws,), f64)
        copy_stream = None
        if host_xi is not None or host_sigma is not None:
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream()
            copy_stream = self._copy_stream
            copy_stream.wait_stream(torch.cuda.current_stream())
        for c0 in range(0, B, Bc):
            bc = min(Bc, B - c0)
            if grouped:
                ids, counts = np.unique(base_host[c0:c0 + bc], return_counts=True)
                ids = np.ascontiguousarray(ids, dtype=np.int32)
                counts = np.ascontiguousarray(counts, dtype=np.int32)
                self._launch('kmpx_inc_adapt_grouped', _lib.ptr(b['s']), _lib.ptr(G), lda, strideA, _lib.ptr(al), _lib.ptr(kap),
                             _lib.ptr(xi_b), _lib.ptr(sg_b), N0, I, O, _lib.ptr(base_of[c0:]), ids.ctypes.data, counts.ctypes.data,
                             len(ids), _lib.ptr(via_s[c0:]), _lib.ptr(via_xi[c0:]), _lib.ptr(via_sigma[c0:]), P, self.tol,
                             self.sigma_f, self.l, self.alpha, bc, _lib.ptr(queries), M, _lib.ptr(ws), nws, _lib.ptr(xi[c0:]),
                             _lib.ptr(sg[c0:]) if sg is not None else None, st)
            else:
                self._launch('kmpx_inc_adapt', _lib.ptr(b['s']), _lib.ptr(G), lda, strideA, _lib.ptr(al), _lib.ptr(kap),
                             _lib.ptr(xi_b), _lib.ptr(sg_b), N0, I, O, _lib.ptr(base_of[c0:]), _lib.ptr(via_s[c0:]),
                             _lib.ptr(via_xi[c0:]), _lib.ptr(via_sigma[c0:]), P, self.tol, self.sigma_f, self.l, self.alpha, bc,
                             _lib.ptr(queries), M, _lib.ptr(ws), nws, _lib.ptr(xi[c0:]), _lib.ptr(sg[c0:]) if sg is not None else None, st)
            if copy_stream is not None:
                done = torch.cuda.Event()
                done.record()
                copy_stream.wait_event(done)
                with torch.cuda.stream(copy_stream):
                    if host_xi is not None:
                        host_xi[c0:c0 + bc].copy_(xi[c0:c0 + bc], non_blocking=True)
                    if host_sigma is not None and sg is not None:
                        host_sigma[c0:c0 + bc].copy_(sg[c0:c0 + bc], non_blocking=True)
        if copy_stream is not None:
            torch.cuda.current_stream().wait_stream(copy_stream)
        return xi, sg, info.clone()

    def check_base_info(self, info):
        """Raises numpy.linalg.LinAlgError when a BASE system of run_incremental was not positive definite (every
        problem of that base would be garbage).  One host read - call it outside timed regions."""
        bad = int((info != 0).sum().item())
        if bad:
            raise np.linalg.LinAlgError(f'{bad} of {info.numel()} base systems are not positive definite')

    # ------------------------------------------------------------------ host-facing call
    def fit_predict(self, base_of, via_s, via_xi, via_sigma, queries, want_cov=True):
        """Host arrays in, host arrays out (the reference-facing call):
        base_of (B,), via_s (B,P) or (B,P,I), via_xi (B,P,O), via_sigma (B,P,O,O), queries (I,M).
        Returns xi (B,O,M), sigma (B,O,O,M)."""
        torch = _lib.require_cuda()
        dev = torch.device('cuda')
        via_s = np.asarray(via_s, dtype=np.float64)
        if via_s.ndim == 2:
            via_s = via_s[:, :, None]
        d_base = torch.as_tensor(np.asarray(base_of, dtype=np.int32), device=dev)
        d_vs = torch.as_tensor(np.ascontiguousarray(via_s), device=dev)
        d_vx = torch.as_tensor(np.ascontiguousarray(via_xi, dtype=np.float64), device=dev)
        d_vg = torch.as_tensor(np.ascontiguousarray(via_sigma, dtype=np.float64), device=dev)
        d_q = torch.as_tensor(np.ascontiguousarray(np.asarray(queries, dtype=np.float64).T), device=dev)
        B, O, M = len(base_of), self._base['O'], d_q.shape[0]
        h_xi = torch.empty((B, O, M), dtype=torch.float64).pin_memory()
        h_sg = torch.empty((B, O, O, M), dtype=torch.float64).pin_memory() if want_cov else None
        _, _, info, _ = self.run_device(d_base, d_vs, d_vx, d_vg, d_q, want_cov, host_xi=h_xi, host_sigma=h_sg)
        torch.cuda.current_stream().synchronize()
        bad = int((info != 0).sum().item())
        if bad:
            raise np.linalg.LinAlgError(f'{bad} of {len(base_of)} systems are not positive definite')
        return h_xi.numpy(), (h_sg.numpy() if h_sg is not None else None)


from .dist import shard_range  # noqa: E402,F401  (re-export: one definition, dist.py)
