"""Find the cfg-3 problems (seed offset = rank) on which the incremental path deviates from the direct one."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib, workloads as wl
from kmp_demos_b200.batch import KMPBatch
mu, sig, _ = wl.load_letter_references()
L = 26; c = wl.CFG1; per = 1024; off = int(sys.argv[1]) if len(sys.argv) > 1 else 3
kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], chunk=2048)
grid = wl.cfg1_kmp_grid(200)
kb.set_base(np.broadcast_to(grid, (L, 1, 200)), mu, sig)
ts, ps, vs = [], [], []
for li in range(L):
    t, p, v = wl.cfg3_via_sets(li, mu[li], per, seed_offset=off); ts.append(t); ps.append(p); vs.append(v)
T = np.concatenate(ts)
base_of = torch.as_tensor(np.repeat(np.arange(L, dtype=np.int32), per), device='cuda')
d_vs = torch.as_tensor(T[:, :, None], device='cuda'); d_vp = torch.as_tensor(np.concatenate(ps), device='cuda'); d_vv = torch.as_tensor(np.concatenate(vs), device='cuda')
q = torch.as_tensor(np.ascontiguousarray(wl.cfg1_query_grid(1000).T), device='cuda')
xd, sd, info, npts = kb.run_device(base_of, d_vs, d_vp, d_vv, q); xd, sd = xd.clone(), sd.clone()
xi, si, _ = kb.run_incremental(base_of, d_vs, d_vp, d_vv, q)
B = len(base_of)
e = (torch.linalg.norm((xi - xd).reshape(B, -1), dim=1) / torch.linalg.norm(xd.reshape(B, -1), dim=1)).cpu().numpy()
bad = np.argsort(-e)[:8]
g = grid[0]
for b in bad:
    d0 = np.abs(g - T[b, 0]); d1 = np.abs(g - T[b, 1])
    print(f'problem {b} letter {b // per}: err {e[b]:.2e} n_pts {int(npts[b])} via_s {T[b].tolist()} nearest {int(d0.argmin())} ({d0.min():.2e}) {int(d1.argmin())} ({d1.min():.2e}) |t0-t1| {abs(T[b,0]-T[b,1]):.3e}')
print('count err > 1e-9:', int((e > 1e-9).sum()))
