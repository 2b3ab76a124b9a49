"""Incremental adaptation vs the direct batched path on cfg-3 problems; timing of both."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib, workloads as wl
from kmp_demos_b200.batch import KMPBatch
mu, sig, _ = wl.load_letter_references()
L = 26; c = wl.CFG1; per = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], chunk=2048)
kb.set_base(np.broadcast_to(wl.cfg1_kmp_grid(200), (L, 1, 200)), mu, sig)
ts, ps, vs = [], [], []
for li in range(L):
    t, p, v = wl.cfg3_via_sets(li, mu[li], per); ts.append(t); ps.append(p); vs.append(v)
base_of = torch.as_tensor(np.repeat(np.arange(L, dtype=np.int32), per), device='cuda')
d_vs = torch.as_tensor(np.concatenate(ts)[:, :, None], device='cuda'); d_vp = torch.as_tensor(np.concatenate(ps), device='cuda'); d_vv = torch.as_tensor(np.concatenate(vs), device='cuda')
q = torch.as_tensor(np.ascontiguousarray(wl.cfg1_query_grid(1000).T), device='cuda')
xd, sd, info, _ = kb.run_device(base_of, d_vs, d_vp, d_vv, q)
xd, sd = xd.clone(), sd.clone()
xi, si, info_b = kb.run_incremental(base_of, d_vs, d_vp, d_vv, q)
torch.cuda.synchronize()
rel = lambda a, b: float(torch.linalg.norm((a - b).reshape(a.shape[0], -1), dim=1).div(torch.linalg.norm(b.reshape(b.shape[0], -1), dim=1)).max())
print(f'{len(base_of)} problems: worst per-problem rel err  mean {rel(xi, xd):.2e}  cov {rel(si, sd):.2e}  base info {int(info_b.abs().sum())}')
for name, fn in (('direct', lambda: kb.run_device(base_of, d_vs, d_vp, d_vv, q)), ('incremental', lambda: kb.run_incremental(base_of, d_vs, d_vp, d_vv, q))):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'{name:12s} {ms:8.2f} ms  {len(base_of) / ms * 1e3:10.0f} problems/s')
kb.events = []
kb.run_incremental(base_of, d_vs, d_vp, d_vv, q); torch.cuda.synchronize()
agg = {}
for n, a, b_ in kb.events: agg[n] = agg.get(n, 0.0) + a.elapsed_time(b_)
print({k: round(v, 3) for k, v in agg.items()})
