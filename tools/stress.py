"""Repeated runs of the multi-role kernels on random ragged batches (looks for timing-dependent hangs / wrong results)."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t0 = time.time(); worst = 0.0; rounds = 0
while time.time() - t0 < (float(sys.argv[2]) if len(sys.argv) > 2 else 60):
    nmax = int(rng.choice([96, 130, 202, 300, 404, 448, 512]))
    batch = int(rng.integers(1, 700))
    lda = nmax + int(rng.choice([0, 2, 4, 8]))
    sizes = rng.integers(max(1, nmax - 40), nmax + 1, batch); sizes[rng.integers(0, batch)] = nmax
    M = rng.normal(size=(nmax, nmax)); S = M @ M.T / nmax + np.eye(nmax)
    A = torch.full((batch, nmax, lda), float('nan'), dtype=torch.float64, device='cuda')
    Sd = torch.as_tensor(np.tril(S), device='cuda')
    for b, n in enumerate(sizes):
        A[b, :n, :n] = Sd[:n, :n]
    n_sys = torch.as_tensor(sizes.astype(np.int32), device='cuda'); info = torch.zeros(batch, dtype=torch.int32, device='cuda')
    _lib.call('kmpx_potrf', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, batch, _lib.ptr(info), None, 0, _lib.stream())
    _lib.call('kmpx_trtri', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, batch, None, 0, _lib.stream())
    torch.cuda.synchronize()
    assert int(info.abs().sum()) == 0
    for b in rng.integers(0, batch, 3):
        n = int(sizes[b]); ref = np.linalg.inv(np.linalg.cholesky(S[:n, :n])); got = np.tril(A[b, :n, :n].cpu().numpy())
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
    rounds += 1
print(f'{rounds} random ragged batches, worst rel err {worst:.2e}')
