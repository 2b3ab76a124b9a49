"""Random ragged batches through the warp-per-problem factor kernels (all variants, with and without the structure flag /
fused solve) against numpy, and random KMPBatch problems (O, I, N, M, via points random) 'warp' vs 'cta' fit and
skipping vs dense predict.  usage: stress_wpp.py [seed] [seconds]"""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib
from kmp_demos_b200.batch import KMPBatch
lib = _lib.load()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 40
t0 = time.time(); worst = 0.0; rounds = 0
while time.time() - t0 < budget / 2:
    nmax = int(rng.choice([64, 96, 130, 202, 300, 404, 448, 512]))
    batch = int(rng.integers(1, 400))
    lda = nmax + int(rng.choice([0, 2, 4, 8]))
    sizes = rng.integers(max(1, nmax - 70), nmax + 1, batch); sizes[rng.integers(0, batch)] = nmax
    M = rng.normal(size=(nmax, nmax)); S = M @ M.T / nmax + np.eye(nmax)
    A = torch.full((batch, nmax, lda), float('nan'), dtype=torch.float64, device='cuda')
    Sd = torch.as_tensor(np.tril(S), device='cuda')
    for b, n in enumerate(sizes):
        A[b, :n, :n] = Sd[:n, :n]
    n_sys = torch.as_tensor(sizes.astype(np.int32), device='cuda'); info = torch.zeros(batch, dtype=torch.int32, device='cuda')
    lib.kmpx_debug_set(int(rng.choice([0, 128, 256])))
    _lib.call('kmpx_potrf_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, batch, _lib.ptr(info), 0, _lib.stream())
    _lib.call('kmpx_trtri_wpp', _lib.ptr(A), lda, nmax * lda, _lib.ptr(n_sys), nmax, batch, 0, _lib.stream())
    torch.cuda.synchronize(); lib.kmpx_debug_set(0)
    assert int(info.abs().sum()) == 0
    for b in rng.integers(0, batch, 3):
        n = int(sizes[b]); ref = np.linalg.inv(np.linalg.cholesky(S[:n, :n])); got = np.tril(A[b, :n, :n].cpu().numpy())
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
        assert torch.isnan(A[b, n:, :]).all()
    rounds += 1
print(f'{rounds} random ragged batches through the warp-per-problem kernels, worst rel err {worst:.2e}', flush=True)
worst_fit = worst_skip = 0.0; rounds = 0
while time.time() - t0 < budget:
    O = int(rng.choice([1, 2])); I = int(rng.choice([1, 2])); N0 = int(rng.integers(40, 230)); L = int(rng.integers(1, 4))
    B = int(rng.integers(1, 40)); P = int(rng.integers(1, 3)); Mq = int(rng.integers(1, 700))
    s_base = np.sort(rng.uniform(0, 2, (L, I, N0)), axis=2)
    xi_base = rng.normal(size=(L, O, N0))
    Mm = rng.normal(size=(L, N0, O, O)) * 0.3
    sg = np.einsum('lnab,lncb->lnac', Mm, Mm) + 0.05 * np.eye(O)
    sigma_base = np.ascontiguousarray(np.transpose(sg, (0, 2, 3, 1)))
    base_of = rng.integers(0, L, B)
    via_s = rng.uniform(-0.2, 2.2, (B, P, I)); via_xi = rng.normal(size=(B, P, O))
    Vm = rng.normal(size=(B, P, O, O)) * 0.1
    via_sigma = np.einsum('bpac,bpdc->bpad', Vm, Vm) + 1e-3 * np.eye(O)
    q = np.sort(rng.uniform(-1.0, 3.0, (I, Mq)), axis=1)
    sigma_f = float(rng.choice([5.0, 50.0, 200.0]))
    res = {}
    for mode in ('cta', 'warp'):
        kb = KMPBatch(l=0.5, alpha=40.0, sigma_f=sigma_f, chunk=int(rng.integers(1, 64)))
        kb.factor_mode = mode
        kb.set_base(s_base, xi_base, sigma_base)
        try:
            res[mode] = kb.fit_predict(base_of, via_s, via_xi, via_sigma, q)
        except ValueError:
            res = None; break          # system larger than the batched path takes
    if res is None:
        continue
    # (means of queries far from every stored point are ~1e-30 in the dense kernel and exactly 0 with the skipping: floor the scale)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-6))
    worst_fit = max(worst_fit, rel(res['warp'][0], res['cta'][0]), rel(res['warp'][1], res['cta'][1]))
    lib.kmpx_predict_skip_threshold(0.0); lib.kmpx_predict_tile_queries(32)       # every term kept, same CTA shape (same summation order)
    dense = kb.fit_predict(base_of, via_s, via_xi, via_sigma, q)
    lib.kmpx_predict_skip_threshold(2.0 ** -68); lib.kmpx_predict_tile_queries(0)
    worst_skip = max(worst_skip, rel(res['warp'][0], dense[0]), rel(res['warp'][1], dense[1]))
    rounds += 1
print(f'{rounds} random KMPBatch problems: warp vs cta fit worst {worst_fit:.2e}, skipping vs dense predict worst {worst_skip:.2e}')
