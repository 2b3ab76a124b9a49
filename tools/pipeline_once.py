"""One pass of the cfg-3 pipeline over 26 x <per> problems in one chunk (for ncu captures of the individual kernels).
usage: pipeline_once.py [per = 204: 5304 problems, three per warp slot of the factor kernels] [kmpx_debug_set flags = 0]"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib, workloads as wl
from kmp_demos_b200.batch import KMPBatch
lib = _lib.load()
mu, sig, _ = wl.load_letter_references()
L = 26; c = wl.CFG1; per = int(sys.argv[1]) if len(sys.argv) > 1 else 204
if len(sys.argv) > 2: lib.kmpx_debug_set(int(sys.argv[2]))
kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], chunk=5328)
kb.set_base(np.broadcast_to(wl.cfg1_kmp_grid(200), (L, 1, 200)), mu, sig)
ts, ps, vs = [], [], []
for li in range(L):
    t, p, v = wl.cfg3_via_sets(li, mu[li], per); ts.append(t); ps.append(p); vs.append(v)
base_of = torch.as_tensor(np.repeat(np.arange(L, dtype=np.int32), per), device='cuda')
d_vs = torch.as_tensor(np.concatenate(ts)[:, :, None], device='cuda'); d_vp = torch.as_tensor(np.concatenate(ps), device='cuda'); d_vv = torch.as_tensor(np.concatenate(vs), device='cuda')
q = torch.as_tensor(np.ascontiguousarray(wl.cfg1_query_grid(1000).T), device='cuda')
for rep in range(2):
    kb.events = []
    kb.run_device(base_of, d_vs, d_vp, d_vv, q); torch.cuda.synchronize()
print({n: round(a.elapsed_time(b), 3) for n, a, b in kb.events})
