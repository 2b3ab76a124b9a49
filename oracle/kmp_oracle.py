"""CPU oracle for the KMP learn-and-adapt hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product path
(``kmp_demos_b200``) never does and fails loudly when the CUDA library is missing.

It restates, in plain numpy, the arithmetic of the reference (lbusellato/KMP_demos) for every row of
SURVEY.md section 8(a).  Two flavours exist for the expensive functions:

* ``*_literal``  - follows the reference's own loop nest (one kernel block per Python iteration), used
  for small cases and as the "reference arm" CPU baseline because it costs what the reference costs;
* vectorised     - same expressions, broadcast; checked against the literal one in tests/.

Parity pinning: the reference has no tests or golden vectors of its own, so
the oracle is pinned against outputs of the reference itself, produced in the authoring container by
``tests/golden/make_golden.py`` (which imports /root/reference through a two-line shim) and committed
as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` holds the comparison.

Third-party arithmetic that the reference calls but does not vendor: scikit-learn (requirements.txt:3,
``>=1.5.2``; 1.9.0 installed) for ``GaussianMixture``/``KMeans``, scipy (``>=1.14.1``; 1.18.1) for
``multivariate_normal``.  Their published algorithms are restated here (k-means++ / Lloyd / EM); the
restatement is pinned against the installed packages by the same golden files.

Citations are ``file:line`` relative to /root/reference unless prefixed ``sklearn:``.
"""
from __future__ import annotations

import math

import numpy as np

REALMIN = np.finfo(np.float64).tiny  # KMP.py:9, GaussianMixtureModel.py:15
FD_DT = 0.001  # finite-difference step of the time-driven kernel, KMP.py:113


# ----------------------------------------------------------------------------------------------
# a1  KMP.__kernel_matrix  (kmp/model/KMP.py:93-128)
# ----------------------------------------------------------------------------------------------
def rbf(s1, s2, sigma_f):
    """exp(-sigma_f * ||s1-s2||^2) with the norm taken first and then squared (KMP.py:109-110)."""
    return np.exp(-sigma_f * np.linalg.norm(np.asarray(s1, float) - np.asarray(s2, float)) ** 2)


def kernel_coeffs(s1, s2, sigma_f, time_driven):
    """The scalar coefficients of one kernel block.

    plain        -> (ktt,)                                        KMP.py:114-116
    time-driven  -> (ktt, ktdt, kdtt, kdtdt), forward differences  KMP.py:117-123
    """
    s1 = np.asarray(s1, float)
    s2 = np.asarray(s2, float)
    ktt = rbf(s1, s2, sigma_f)
    if not time_driven:
        return (ktt,)
    a = rbf(s1, s2 + FD_DT, sigma_f)
    b = rbf(s1 + FD_DT, s2, sigma_f)
    c = rbf(s1 + FD_DT, s2 + FD_DT, sigma_f)
    ktdt = (a - ktt) / FD_DT
    kdtt = (b - ktt) / FD_DT
    kdtdt = (c - a - b + ktt) / FD_DT ** 2
    return (ktt, ktdt, kdtt, kdtdt)


def kernel_block(s1, s2, sigma_f, O, time_driven):
    """(O,O) kernel block: k*I_O, or [[ktt,ktdt],[kdtt,kdtdt]] (x) I_{O/2}  (KMP.py:115-126)."""
    co = kernel_coeffs(s1, s2, sigma_f, time_driven)
    if not time_driven:
        return co[0] * np.eye(O)
    P = O // 2
    eye = np.eye(P)
    return np.block([[co[0] * eye, co[1] * eye], [co[2] * eye, co[3] * eye]])


# ----------------------------------------------------------------------------------------------
# a2  KMP.fit  (kmp/model/KMP.py:130-158)
# ----------------------------------------------------------------------------------------------
def kmp_matrix_literal(s, sigma, lam, sigma_f, time_driven):
    """A[iO:(i+1)O, jO:(j+1)O] = kernel(s_i,s_j) + delta_ij*lam*Sigma_i, N^2 Python iterations
    (KMP.py:146-156).  s:(I,N)  sigma:(O,O,N)."""
    O, N = sigma.shape[0], sigma.shape[2]
    A = np.zeros((N * O, N * O))
    for i in range(N):
        for j in range(N):
            A[i * O:(i + 1) * O, j * O:(j + 1) * O] = kernel_block(s[:, i], s[:, j], sigma_f, O, time_driven)
            if i == j:
                A[j * O:(j + 1) * O, i * O:(i + 1) * O] += lam * sigma[:, :, i]
    return A


def _coeff_matrices(sa, sb, sigma_f, time_driven):
    """Broadcast form of kernel_coeffs for all pairs: sa:(I,Na), sb:(I,Nb) -> tuple of (Na,Nb).

    Uses exactly the reference's expression order (norm -> square -> scale -> exp) so that the result
    equals the literal path bit for bit on the same numpy build (SURVEY.md 8a1 [probed])."""
    def k(x, y):
        d = x[:, :, None] - y[:, None, :]
        r = np.sqrt(np.sum(d * d, axis=0)) if d.shape[0] > 1 else np.abs(d[0])
        return np.exp(-sigma_f * r ** 2)
    ktt = k(sa, sb)
    if not time_driven:
        return (ktt,)
    a = k(sa, sb + FD_DT)
    b = k(sa + FD_DT, sb)
    c = k(sa + FD_DT, sb + FD_DT)
    return (ktt, (a - ktt) / FD_DT, (b - ktt) / FD_DT, (c - a - b + ktt) / FD_DT ** 2)


def _expand_blocks(co, O, time_driven):
    """(Na,Nb) coefficient matrices -> (Na*O, Nb*O) point-major block matrix."""
    if not time_driven:
        return np.kron(co[0], np.eye(O))
    P = O // 2
    Na, Nb = co[0].shape
    out = np.zeros((Na, O, Nb, O))
    eye = np.eye(P)
    out[:, :P, :, :P] = co[0][:, None, :, None] * eye[None, :, None, :]
    out[:, :P, :, P:] = co[1][:, None, :, None] * eye[None, :, None, :]
    out[:, P:, :, :P] = co[2][:, None, :, None] * eye[None, :, None, :]
    out[:, P:, :, P:] = co[3][:, None, :, None] * eye[None, :, None, :]
    return out.reshape(Na * O, Nb * O)


def kmp_matrix(s, sigma, lam, sigma_f, time_driven):
    """Vectorised kmp_matrix_literal (same values)."""
    O, N = sigma.shape[0], sigma.shape[2]
    A = _expand_blocks(_coeff_matrices(s, s, sigma_f, time_driven), O, time_driven)
    for i in range(N):
        A[i * O:(i + 1) * O, i * O:(i + 1) * O] += lam * sigma[:, :, i]
    return A


def kmp_fit(s, sigma, lam, sigma_f, time_driven, literal=False):
    """Returns the reference's ``_estimator`` = inv(A) (LU inverse, KMP.py:157)."""
    A = (kmp_matrix_literal if literal else kmp_matrix)(s, sigma, lam, sigma_f, time_driven)
    return np.linalg.inv(A)


# ----------------------------------------------------------------------------------------------
# a3  KMP.predict  (kmp/model/KMP.py:160-194)
# ----------------------------------------------------------------------------------------------
def kmp_predict_literal(estimator, s_train, xi_train, s_query, alpha, sigma_f, time_driven):
    """Per-query loop of the reference (KMP.py:178-190): k (O x NO), Y point-major, k@inv@Y,
    alpha*(kernel(s*,s*) - k@inv@k.T)."""
    O, N = xi_train.shape
    M = s_query.shape[1]
    xi = np.zeros((O, M))
    sig = np.zeros((O, O, M))
    for j in range(M):
        k = np.zeros((O, N * O))
        Y = np.zeros(N * O)
        for i in range(N):
            k[:, i * O:(i + 1) * O] = kernel_block(s_query[:, j], s_train[:, i], sigma_f, O, time_driven)
            for h in range(O):
                Y[i * O + h] = xi_train[h, i]
        xi[:, j] = k @ estimator @ Y
        sig[:, :, j] = alpha * (kernel_block(s_query[:, j], s_query[:, j], sigma_f, O, time_driven)
                                - k @ estimator @ k.T)
    return xi, sig


def kmp_predict(estimator, s_train, xi_train, s_query, alpha, sigma_f, time_driven):
    """Vectorised kmp_predict_literal: same operand grouping ((k@inv)@Y and (k@inv)@k.T)."""
    O, N = xi_train.shape
    M = s_query.shape[1]
    Kq = _expand_blocks(_coeff_matrices(s_query, s_train, sigma_f, time_driven), O, time_driven)  # (MO, NO)
    Y = xi_train.T.reshape(-1)
    W = Kq @ estimator                                   # (MO, NO)
    xi = (W @ Y).reshape(M, O).T.copy()
    W3 = W.reshape(M, O, N * O)
    K3 = Kq.reshape(M, O, N * O)
    quad = np.einsum('mac,mbc->mab', W3, K3)
    kss = np.stack([kernel_block(s_query[:, j], s_query[:, j], sigma_f, O, time_driven) for j in range(M)]) \
        if time_driven else np.broadcast_to(np.eye(O), (M, O, O))
    sig = alpha * (kss - quad)
    return xi, np.ascontiguousarray(np.transpose(sig, (1, 2, 0)))


# ----------------------------------------------------------------------------------------------
# a4  KMP.set_waypoint  (kmp/model/KMP.py:57-91)
# ----------------------------------------------------------------------------------------------
def kmp_set_waypoint(s_db, xi_db, sigma_db, s, xi, sigma, tol):
    """Nearest stored input by L2, strict ``< tol`` => overwrite, else append (KMP.py:72-89).
    Returns the new (s, xi, sigma) database; the caller refits (KMP.py:91).
    The scan covers only the points present at the last fit (``self.N`` is not updated between
    via-points of one call, KMP.py:74), so a point appended earlier in the same call is never matched."""
    s_db, xi_db, sigma_db = s_db.copy(), xi_db.copy(), sigma_db.copy()
    n_at_entry = s_db.shape[1]
    for j in range(len(s)):
        min_dist, idx = np.inf, None
        for i in range(n_at_entry):
            dist = np.linalg.norm(s_db[:, i] - s[j])
            if dist < min_dist:
                min_dist, idx = dist, i
        if min_dist < tol:
            s_db[:, idx] = s[j]
            xi_db[:, idx] = xi[j]
            sigma_db[:, :, idx] = sigma[j]
        else:
            s_db = np.concatenate((s_db, np.array(s[j]).reshape(1, -1)), axis=1)
            xi_db = np.concatenate((xi_db, xi[j].T), axis=1)
            sigma_db = np.concatenate((sigma_db, np.expand_dims(sigma[j], 2)), axis=2)
    return s_db, xi_db, sigma_db


# ----------------------------------------------------------------------------------------------
# a6  GaussianMixtureModel.predict = GMR  (kmp/mixture/GaussianMixtureModel.py:59-107)
# ----------------------------------------------------------------------------------------------
def _mvn_pdf(x, mu, cov):
    """scipy.stats.multivariate_normal(mu,cov).pdf(x) restated (eigh-based pseudo-inverse / log-pdet,
    scipy ``_PSD`` + ``_logpdf``)."""
    cov = np.atleast_2d(cov)
    w, v = np.linalg.eigh(cov)
    eps = 1e6 * np.finfo(float).eps * np.max(np.abs(w))
    keep = w > eps
    d = np.where(keep, w, 1.0)
    U = v[:, keep] / np.sqrt(d[keep])
    log_pdet = np.sum(np.log(d[keep]))
    dev = np.atleast_1d(x) - np.atleast_1d(mu)
    maha = np.sum((dev @ U) ** 2)
    rank = int(keep.sum())
    return math.exp(-0.5 * (rank * math.log(2 * math.pi) + log_pdet + maha))


def gmr_literal(priors, means, covs, data, reg):
    """The reference's triple loop (GMM.py:83-105).  means:(d,K) covs:(d,d,K) data:(I,M)."""
    I, M = data.shape
    d, K = means.shape
    O = d - I
    out_mu = np.zeros((O, M))
    out_cov = np.zeros((O, O, M))
    mu_tmp = np.zeros((O, K))
    H = np.zeros((K, M))
    for t in range(M):
        for i in range(K):
            H[i, t] = priors[i] * _mvn_pdf(data[:, t], means[:I, i], covs[:I, :I, i])
        H[:, t] /= np.sum(H[:, t] + REALMIN)
        for i in range(K):
            g = covs[I:, :I, i] @ np.linalg.inv(covs[:I, :I, i])
            mu_tmp[:, i] = means[I:, i] + g @ (data[:, t] - means[:I, i])
            out_mu[:, t] += H[i, t] * mu_tmp[:, i]
        for i in range(K):
            sc = covs[I:, I:, i] - covs[I:, :I, i] @ np.linalg.inv(covs[:I, :I, i]) @ covs[:I, I:, i]
            out_cov[:, :, t] += H[i, t] * (sc + np.outer(mu_tmp[:, i], mu_tmp[:, i]))
        out_cov[:, :, t] += np.eye(O) * reg - np.outer(out_mu[:, t], out_mu[:, t])
    return out_mu, out_cov


def gmr(priors, means, covs, data, reg):
    """Closed-form GMR with per-component constants hoisted (Cholesky log-pdf)."""
    I, M = data.shape
    d, K = means.shape
    O = d - I
    G = np.zeros((K, O, I))
    Sc = np.zeros((K, O, O))
    logc = np.zeros(K)
    Linv = np.zeros((K, I, I))
    for i in range(K):
        Sii = covs[:I, :I, i]
        Sinv = np.linalg.inv(Sii)
        G[i] = covs[I:, :I, i] @ Sinv
        Sc[i] = covs[I:, I:, i] - covs[I:, :I, i] @ Sinv @ covs[:I, I:, i]
        L = np.linalg.cholesky(Sii)
        Linv[i] = np.linalg.inv(L)
        logc[i] = -0.5 * I * math.log(2 * math.pi) - np.sum(np.log(np.diag(L)))
    out_mu = np.zeros((O, M))
    out_cov = np.zeros((O, O, M))
    for t in range(M):
        x = data[:, t]
        h = np.zeros(K)
        mu_c = np.zeros((K, O))
        for i in range(K):
            dev = x - means[:I, i]
            z = Linv[i] @ dev
            h[i] = priors[i] * math.exp(logc[i] - 0.5 * float(z @ z))
            mu_c[i] = means[I:, i] + G[i] @ dev
        h = h / np.sum(h + REALMIN)
        m = h @ mu_c
        c = np.zeros((O, O))
        for i in range(K):
            c += h[i] * (Sc[i] + np.outer(mu_c[i], mu_c[i]))
        out_mu[:, t] = m
        out_cov[:, :, t] = c + np.eye(O) * reg - np.outer(m, m)
    return out_mu, out_cov


# ----------------------------------------------------------------------------------------------
# a5  GaussianMixtureModel.fit -> sklearn GaussianMixture  (GMM.py:40-41,52-57)
#     sklearn:cluster/_kmeans.py:180-278 (k-means++), :630-758 (Lloyd), :1463-1565 (fit),
#     sklearn:cluster/_k_means_lloyd.pyx:23-230, sklearn:cluster/_k_means_common.pyx:167-311,
#     sklearn:mixture/_base.py:104-128,172-312, sklearn:mixture/_gaussian_mixture.py:168-197,282-385,490-553
# ----------------------------------------------------------------------------------------------
def _sq_dists_expanded(C, X, x_sq):
    """sklearn ``_euclidean_distances(C, X, Y_norm_squared=x_sq, squared=True)``: -2 C.X^T + |c|^2 + |x|^2,
    clipped at 0 (sklearn:metrics/pairwise.py)."""
    c_sq = np.einsum('ij,ij->i', C, C)[:, None]
    D = -2.0 * (C @ X.T)
    D += c_sq
    D += x_sq[None, :]
    np.maximum(D, 0, out=D)
    return D


def kmeans_plusplus(X, K, x_sq, rs):
    """Greedy k-means++ (sklearn:cluster/_kmeans.py:180-278) with unit sample weights, driven by the
    host RandomState ``rs`` exactly as sklearn drives it."""
    n = X.shape[0]
    n_trials = 2 + int(np.log(K))
    w = np.ones(n)
    centers = np.empty((K, X.shape[1]))
    idx = np.full(K, -1, dtype=int)
    cid = rs.choice(n, p=w / w.sum())
    centers[0] = X[cid]
    idx[0] = cid
    closest = _sq_dists_expanded(centers[0, None], X, x_sq)
    pot = closest @ w
    for c in range(1, K):
        rand_vals = rs.uniform(size=n_trials) * pot
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)
        np.clip(cand, None, closest.size - 1, out=cand)
        dc = _sq_dists_expanded(X[cand], X, x_sq)
        np.minimum(closest, dc, out=dc)
        cpot = dc @ w.reshape(-1, 1)
        best = int(np.argmin(cpot))
        pot = cpot[best]
        closest = dc[best]
        centers[c] = X[cand[best]]
        idx[c] = cand[best]
    return centers, idx


def _lloyd_iter(X, centers_old, update=True):
    """One ``lloyd_iter_chunked_dense`` (sklearn:cluster/_k_means_lloyd.pyx:23-230): labels by
    argmin_j(|c_j|^2 - 2 x.c_j), first minimum wins; unit weights."""
    K = centers_old.shape[0]
    c_sq = np.einsum('ij,ij->i', centers_old, centers_old)
    pd = c_sq[None, :] - 2.0 * (X @ centers_old.T)
    labels = np.argmin(pd, axis=1).astype(np.int32)
    if not update:
        return labels, None, None
    wsum = np.bincount(labels, minlength=K).astype(float)
    new = np.zeros_like(centers_old)
    np.add.at(new, labels, X)
    empty = np.where(wsum == 0)[0]
    if empty.size:  # sklearn:cluster/_k_means_common.pyx:167-211
        dist = ((X - centers_old[labels]) ** 2).sum(axis=1)
        far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
        if np.max(dist) != 0:
            for e, f in zip(empty, far):
                old = labels[f]
                new[old] -= X[f]
                new[e] = X[f]
                wsum[e] = 1.0
                wsum[old] -= 1.0
    amax = int(np.argmax(wsum))
    for j in range(K):  # _average_centers, _k_means_common.pyx:274-295
        if wsum[j] > 0:
            new[j] *= 1.0 / wsum[j]
        else:
            new[j] = new[amax]
    shift = np.sqrt(((new - centers_old) ** 2).sum(axis=1))
    return labels, new, shift


def sklearn_kmeans_labels(X, K, rs, max_iter=300, tol=1e-4):
    """``KMeans(n_clusters=K, n_init=1, random_state=rs).fit(X).labels_`` restated
    (sklearn:cluster/_kmeans.py:1463-1565 and :630-758)."""
    X = np.array(X, dtype=float, order='C')
    X = X - X.mean(axis=0)
    x_sq = np.einsum('ij,ij->i', X, X)
    tol_abs = np.mean(np.var(X, axis=0)) * tol
    centers, _ = kmeans_plusplus(X, K, x_sq, rs)
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    n_iter = 0
    for it in range(max_iter):
        labels, new, shift = _lloyd_iter(X, centers)
        centers = new
        n_iter = it + 1
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if (shift ** 2).sum() <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels, _, _ = _lloyd_iter(X, centers, update=False)
    return labels, n_iter


def gmm_m_step(X, resp, reg):
    """``_estimate_gaussian_parameters`` for 'full' (sklearn:mixture/_gaussian_mixture.py:168-197,282-320)."""
    nk = resp.sum(axis=0) + 10 * np.finfo(float).eps
    means = (resp.T @ X) / nk[:, None]
    K, d = means.shape
    covs = np.empty((K, d, d))
    for k in range(K):
        diff = X - means[k]
        covs[k] = ((resp[:, k] * diff.T) @ diff) / nk[k]
        covs[k].flat[::d + 1] += reg
    return nk, means, covs


def gmm_precisions_chol(covs):
    """``_compute_precision_cholesky`` 'full' (sklearn:mixture/_gaussian_mixture.py:358-370): P_k = (chol S_k)^-T."""
    K, d, _ = covs.shape
    P = np.empty_like(covs)
    for k in range(K):
        L = np.linalg.cholesky(covs[k])
        P[k] = np.linalg.solve(L, np.eye(d)).T
    return P


def gmm_e_step(X, weights, means, P):
    """``_estimate_log_prob_resp`` (sklearn:mixture/_base.py:560-590, _gaussian_mixture.py:490-553):
    returns (per-sample log p(x), log responsibilities)."""
    n, d = X.shape
    K = means.shape[0]
    log_det = np.array([np.sum(np.log(np.diag(P[k]))) for k in range(K)])
    lp = np.empty((n, K))
    for k in range(K):
        y = (X - means[k]) @ P[k]
        lp[:, k] = np.sum(y * y, axis=1)
    wlp = -0.5 * (d * math.log(2 * math.pi) + lp) + log_det + np.log(weights)
    mx = wlp.max(axis=1, keepdims=True)
    lse = mx[:, 0] + np.log(np.sum(np.exp(wlp - mx), axis=1))
    with np.errstate(under='ignore'):
        log_resp = wlp - lse[:, None]
    return lse, log_resp


def gmm_fit(X, K, reg, seed=420, tol=1e-3, max_iter=100, init_labels=None):
    """``GaussianMixture(n_components=K, reg_covar=reg, random_state=seed).fit(X)`` restated
    (sklearn:mixture/_base.py:172-312).  Returns a dict with weights (K,), means (K,d),
    covariances (K,d,d), n_iter, lower_bound, converged, init_labels."""
    X = np.asarray(X, dtype=float)
    n = X.shape[0]
    if init_labels is None:
        rs = np.random.RandomState(seed)
        init_labels, _ = sklearn_kmeans_labels(X, K, rs)
    resp = np.zeros((n, K))
    resp[np.arange(n), init_labels] = 1
    nk, means, covs = gmm_m_step(X, resp, reg)
    weights = nk / n
    P = gmm_precisions_chol(covs)
    lower = -np.inf
    converged = False
    n_iter = 0
    for it in range(1, max_iter + 1):
        prev = lower
        lse, log_resp = gmm_e_step(X, weights, means, P)
        nk, means, covs = gmm_m_step(X, np.exp(log_resp), reg)
        weights = nk / nk.sum()
        P = gmm_precisions_chol(covs)
        lower = float(np.mean(lse))
        n_iter = it
        if abs(lower - prev) < tol:
            converged = True
            break
    return dict(weights=weights, means=means, covariances=covs, precisions_chol=P, n_iter=n_iter,
                lower_bound=lower, converged=converged, init_labels=np.asarray(init_labels))


def gmm_score_samples(X, weights, means, covs):
    lse, _ = gmm_e_step(np.asarray(X, float), weights, means, gmm_precisions_chol(covs))
    return lse


def gmm_bic(X, weights, means, covs):
    """sklearn ``bic``: -2*score*n + n_params*log(n) (GMM.py:109-122)."""
    n, d = X.shape
    K = means.shape[0]
    n_params = int(K * d * (d + 1) / 2.0 + d * K + K - 1)
    return -2 * float(np.mean(gmm_score_samples(X, weights, means, covs))) * n + n_params * math.log(n)


# ----------------------------------------------------------------------------------------------
# a8/a9  kmp/cluster/Cluster.py
# ----------------------------------------------------------------------------------------------
def uniform_labels(n_total, n_components, n_demos):
    """Uniform.fit_predict (Cluster.py:33-47)."""
    m = int(n_total / n_demos)
    return np.tile((np.arange(m) * n_components / m).astype(int), n_demos)


def ref_kmeans(X, K, n_iter, tol, init_idx):
    """kmp.cluster.KMeans.fit_predict (Cluster.py:74-102) with the initial indices given (the reference
    draws them from the global numpy RNG, Cluster.py:88).  X:(d,n).  Returns (labels, iterations)."""
    centroids = X[:, init_idx].T
    it_done = n_iter
    with np.errstate(invalid='ignore', divide='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for iteration in range(n_iter):
                distances = np.array([np.linalg.norm(X.T - c, axis=1) for c in centroids])
                labels = np.argmin(distances, axis=0)
                new_centroids = np.array([X[:, labels == i].mean(axis=1) for i in range(K)])
                if np.allclose(centroids, new_centroids, rtol=tol):
                    it_done = iteration + 1
                    break
                centroids = new_centroids
    return np.argmin(distances, axis=0), it_done


# ----------------------------------------------------------------------------------------------
# a10-a12  kmp/types/quaternion.py   (arrays are scalar-first [v, u0, u1, u2])
# ----------------------------------------------------------------------------------------------
def quat_normalize(q):
    """The constructor's normalisation (quaternion.py:18-21,121-129)."""
    q = np.asarray(q, float)
    return q / np.linalg.norm(q)


def quat_exp(w):
    """quaternion.exp (quaternion.py:60-80); identity when allclose(w,0)."""
    w = np.asarray(w, float)
    nrm = np.linalg.norm(w)
    q = np.array([1.0, 0.0, 0.0, 0.0])
    if not np.allclose(w, np.zeros_like(w)):
        q = np.concatenate(([np.cos(nrm)], np.sin(nrm) * w / nrm))
    return quat_normalize(q)


def quat_log(q):
    """quaternion.log (quaternion.py:82-96) applied to an already-normalised quaternion array."""
    q = np.asarray(q, float)
    nrm = np.linalg.norm(q)
    u = np.sign(q[0]) * q[1:] / nrm
    v = np.sign(q[0]) * q[0] / nrm
    if not np.allclose(u, np.zeros_like(u)):
        return np.arccos(v) * u / np.linalg.norm(u)
    return np.zeros_like(u)


def quat_mul(p, q):
    """Hamilton product followed by the constructor's normalisation (quaternion.py:182-202)."""
    p = np.asarray(p, float)
    q = np.asarray(q, float)
    u = p[0] * q[1:] + q[0] * p[1:] + np.cross(p[1:], q[1:])
    v = p[0] * q[0] - p[1:] @ q[1:]
    return quat_normalize(np.concatenate(([v], u)))


def quat_conj(q):
    """~q, re-normalised (quaternion.py:219-227)."""
    q = np.asarray(q, float)
    return quat_normalize(np.concatenate(([q[0]], -q[1:])))


def quat_from_rotation_vector(r):
    """quaternion.from_rotation_vector incl. the abs(cos) quirk (quaternion.py:23-42)."""
    r = np.asarray(r, float)
    with np.errstate(invalid='ignore', divide='ignore'):
        ang = np.linalg.norm(r)
        axis = r / ang
        return quat_normalize(np.concatenate(([np.abs(np.cos(ang / 2))], axis * np.sin(ang / 2))))


# ----------------------------------------------------------------------------------------------
# f1  Incremental adaptation: the same numbers as set_waypoint + fit + predict of a2-a4, obtained from
#     the BASE solution by block-inverse identities.  This is the algebra the device kernels of csrc/incremental.cu
#     implement; it is pinned against the direct restatement above (tests/test_oracle_golden.py).
# ----------------------------------------------------------------------------------------------
def kmp_incremental(s_b, xi_b, sigma_b, via_s, via_xi, via_sigma, s_query, lam, alpha, sigma_f, tol):
    """Plain kernel, I = 1.  s_b (1,N0), xi_b (O,N0), sigma_b (O,O,N0): base trajectory; via_s list of P floats,
    via_xi list of (1,O), via_sigma list of (O,O) as KMP.set_waypoint takes them (KMP.py:57); s_query (1,M).
    A via point closer than ``tol`` to a stored input (first minimum, KMP.py:75-80) DELETES that stored point; every
    via point is then APPENDED.  Point-major unknown order (KMP.py:151)."""
    N0 = s_b.shape[1]
    O = xi_b.shape[0]
    M = s_query.shape[1]
    A_b = kmp_matrix(s_b, sigma_b, lam, sigma_f, False)
    G = np.linalg.inv(A_b)
    kap = np.exp(-sigma_f * (s_b[0][:, None] - s_query[0][None, :]) ** 2)          # (N0, M)
    alpha_b = G @ xi_b.T.reshape(-1)
    eye = np.eye(O)
    dele, app, owner = [], [], {}
    s_work = s_b[0].copy()                      # the scan sees earlier in-place overwrites of the same call (KMP.py:80-84)
    for j in range(len(via_s)):
        d = np.abs(s_work - via_s[j])
        idx = int(np.argmin(d))
        if d[idx] < tol:
            if idx in owner:
                app.remove(owner[idx])          # a later via point on the same stored point supersedes the earlier one
            else:
                dele.append(idx)
            owner[idx] = j
            s_work[idx] = via_s[j]
        app.append(j)
    r1, r2 = O * len(dele), O * len(app)
    didx = np.array([i * O + a for i in dele for a in range(O)], dtype=int)
    H = np.linalg.inv(G[np.ix_(didx, didx)]) if r1 else np.zeros((0, 0))
    C = np.zeros((N0 * O, r2))
    D = np.zeros((r2, r2))
    ya = np.zeros(r2)
    for k, j in enumerate(app):
        kc = np.exp(-sigma_f * (s_b[0] - via_s[j]) ** 2)
        kc[dele] = 0.0
        C[:, k * O:(k + 1) * O] = np.kron(kc[:, None], eye)
        ya[k * O:(k + 1) * O] = np.asarray(via_xi[j]).reshape(-1)
        for k2, j2 in enumerate(app):
            D[k * O:(k + 1) * O, k2 * O:(k2 + 1) * O] = np.exp(-sigma_f * (via_s[j] - via_s[j2]) ** 2) * eye
        D[k * O:(k + 1) * O, k * O:(k + 1) * O] += lam * np.asarray(via_sigma[j])
    GC = G @ C
    GCd = GC[didx]
    N1 = H @ GCd
    Sinv = np.linalg.inv(D - (C.T @ GC - GCd.T @ H @ GCd))
    a_p = -H @ alpha_b[didx]
    a_z = Sinv @ (ya - (C.T @ alpha_b - GCd.T @ H @ alpha_b[didx]))
    W = np.concatenate([G[:, didx], GC], axis=1)
    xi = np.empty((O, M))
    sg = np.empty((O, O, M))
    for m in range(M):
        v = np.kron(kap[:, m][None, :], eye)                                       # k_b(s*)  (O, n0)
        u = v @ W
        p, q = u[:, :r1], u[:, r1:]
        ka = np.zeros((O, r2))
        for k, j in enumerate(app):
            ka[:, k * O:(k + 1) * O] = np.exp(-sigma_f * (s_query[0, m] - via_s[j]) ** 2) * eye
        z = ka - q + p @ N1
        quad = v @ G @ v.T - p @ H @ p.T + z @ Sinv @ z.T
        sg[:, :, m] = alpha * (eye - quad)
        xi[:, m] = v @ alpha_b + p @ a_p + z @ a_z
    return xi, sg
