"""Concrete inputs for the BASELINE.json configs.  Pure numpy, no device code,
no oracle import: shared by tests/, bench.py and tests/golden/make_golden.py so that every arm sees
the same seeded inputs."""
from __future__ import annotations

import os

import numpy as np

LETTERS = [chr(ord('A') + i) for i in range(26)]
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# cfg-1 / cfg-3 hyper-parameters (kmp/demos.py:142,158 of the reference)
CFG1 = dict(K=8, reg=1e-6, lam=0.5, alpha=40.0, sigma_f=200.0, time_driven=False, N=200, O=2, M=1000)


def cfg1_design_matrix(pos_demos):
    """pos_demos: (H, 2, T) -> X (H*T, 3) = [t, x, y] with t = 0.01*tile(1..T) (demos.py:136-138)."""
    H, _, T = pos_demos.shape
    pos = np.concatenate([pos_demos[h] for h in range(H)], axis=1)
    t = 0.01 * np.tile(np.linspace(1, T, T), H).reshape(1, -1)
    return np.vstack((t, pos)).T.copy()


def cfg1_gmr_grid(T=200):
    """GMR inputs 0.00..(T-1)*0.01 (demos.py:139) - the reference's off-by-one is kept on purpose."""
    return (0.01 * np.arange(T)).reshape(1, -1)


def cfg1_kmp_grid(T=200):
    """KMP time axis 0.01..T*0.01 (demos.py:155)."""
    return (0.01 * np.arange(1, T + 1)).reshape(1, -1)


def cfg1_query_grid(M=1000):
    return np.linspace(0.01, 2.00, M).reshape(1, -1)


def cfg3_via_sets(letter_idx, mu_ref, n_sets, n_via=2, seed_offset=0):
    """Via-point sets for one letter (SURVEY.md 8d cfg 3).

    For each set: ``n_via`` distinct times; each is on the 0.01 grid with probability 1/2 (the
    reference's replace branch, KMP.py:80-84) or uniform off-grid (append branch, KMP.py:85-89);
    p = mu(t) + N(0, 1.5^2 I); var = 1e-4 I (demos.py:91).
    Returns t:(n_sets,n_via)  p:(n_sets,n_via,O)  var:(n_sets,n_via,O,O)."""
    O, T = mu_ref.shape
    grid = 0.01 * np.arange(1, T + 1)
    t = np.empty((n_sets, n_via))
    noise = np.empty((n_sets, n_via, O))
    for s in range(n_sets):
        rng = np.random.default_rng((int(letter_idx), int(seed_offset), s))   # one stream per set
        while True:
            on = rng.random(n_via) < 0.5
            ti = np.where(on, grid[rng.integers(0, T, n_via)], rng.uniform(0.01, 0.01 * T, n_via))
            if len(np.unique(np.round(ti, 6))) == n_via:
                break
        t[s] = ti
        noise[s] = rng.normal(0.0, 1.5, size=(n_via, O))
    mu_t = np.stack([np.interp(t.reshape(-1), grid, mu_ref[o]) for o in range(O)], axis=-1).reshape(n_sets, n_via, O)
    p = mu_t + noise
    var = np.broadcast_to(1e-4 * np.eye(O), (n_sets, n_via, O, O)).copy()
    return t, p, var


def cfg4_problem(N=8192, seed=0):
    """Synthetic single large KMP (SURVEY.md 8d cfg 4): t=linspace(0,2,N), mu=8(sin,cos)(2 pi t),
    Sigma_i = R(th) diag(.05+.5u, .05+.5v) R(th)^T."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 2.0, N).reshape(1, -1)
    mu = 8.0 * np.vstack((np.sin(2 * np.pi * t[0]), np.cos(2 * np.pi * t[0])))
    u, v, th = rng.random(N), rng.random(N), rng.random(N) * 2 * np.pi
    c, s = np.cos(th), np.sin(th)
    d0, d1 = 0.05 + 0.5 * u, 0.05 + 0.5 * v
    sig = np.empty((2, 2, N))
    sig[0, 0] = c * c * d0 + s * s * d1
    sig[0, 1] = sig[1, 0] = c * s * (d0 - d1)
    sig[1, 1] = s * s * d0 + c * c * d1
    return t, mu, sig


def cfg5_mixture(K=64, d=7, seed=0):
    """The ground-truth mixture of cfg 5: means (K,d) along a smooth random curve, factors A (K,d,d) of random SPD
    covariances with eigenvalues in [1e-3,1e-1]; also returns the generator positioned after these draws."""
    rng = np.random.default_rng(seed)
    knots = np.linspace(0.0, 1.0, K)
    freq = rng.uniform(0.5, 2.0, size=d)
    phase = rng.uniform(0, 2 * np.pi, size=d)
    means = np.sin(2 * np.pi * knots[:, None] * freq[None, :] + phase[None, :])
    means[:, 0] = knots * 1.5
    A = np.empty((K, d, d))
    for k in range(K):
        Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
        ev = 10.0 ** rng.uniform(-3, -1, size=d)
        A[k] = Q * np.sqrt(ev)
    return means, A, rng


def cfg5_samples(n, K=64, d=7, seed=0):
    """Synthetic GMM EM data (SURVEY.md 8d cfg 5): samples of a K-component mixture whose means follow
    a smooth random curve in d dims and whose covariances are random SPD with eigenvalues in [1e-3,1e-1]."""
    means, A, rng = cfg5_mixture(K, d, seed)
    comp = rng.integers(0, K, size=n)
    z = rng.normal(size=(n, d))
    X = means[comp] + np.einsum('nij,nj->ni', A[comp], z)
    return np.ascontiguousarray(X), means, A


def synthetic_letters(n_letters=26, T=200, seed=7):
    """Fallback trajectories of the cfg-1 shape when tests/golden/letters_ref.npz is absent:
    mu (L,2,T) smooth Lissajous-like curves in [-10,10], sigma (L,2,2,T) SPD."""
    rng = np.random.default_rng(seed)
    tt = np.linspace(0, 1, T)
    mu = np.empty((n_letters, 2, T))
    sig = np.empty((n_letters, 2, 2, T))
    for l in range(n_letters):
        a, b = rng.uniform(0.5, 2.5, 2)
        ph = rng.uniform(0, 2 * np.pi, 2)
        mu[l, 0] = 8 * np.sin(2 * np.pi * a * tt + ph[0])
        mu[l, 1] = 8 * np.cos(2 * np.pi * b * tt + ph[1])
        th = 2 * np.pi * (tt + rng.random())
        c, s = np.cos(th), np.sin(th)
        d0 = 0.05 + 0.5 * (0.5 + 0.5 * np.sin(7 * tt + l))
        d1 = 0.05 + 0.5 * (0.5 + 0.5 * np.cos(5 * tt + l))
        sig[l, 0, 0] = c * c * d0 + s * s * d1
        sig[l, 0, 1] = sig[l, 1, 0] = c * s * (d0 - d1)
        sig[l, 1, 1] = s * s * d0 + c * c * d1
    return mu, sig


def load_letter_references():
    """(mu (26,2,200), sigma (26,2,2,200), source) - the reference's own GMM+GMR output for every letter
    when the golden fixture is present, else synthetic trajectories of the same shape."""
    path = os.path.join(GOLDEN_DIR, 'letters_ref.npz')
    if os.path.isfile(path):
        z = np.load(path)
        return z['gmr_mu'], z['gmr_sigma'], 'tests/golden/letters_ref.npz (reference GMM+GMR of 2Dletters)'
    mu, sig = synthetic_letters()
    return mu, sig, 'synthetic letters (fixture absent)'
