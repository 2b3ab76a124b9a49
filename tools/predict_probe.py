import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib, workloads as wl
from kmp_demos_b200.batch import KMPBatch
lib = _lib.load()
mu, sig, _ = wl.load_letter_references()
L = 26; c = wl.CFG1
kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], chunk=2048)
kb.set_base(np.broadcast_to(wl.cfg1_kmp_grid(200), (L, 1, 200)), mu, sig)
ts, ps, vs = [], [], []
for li in range(L):
    t, p, v = wl.cfg3_via_sets(li, mu[li], 64); ts.append(t); ps.append(p); vs.append(v)
base_of = torch.as_tensor(np.repeat(np.arange(L, dtype=np.int32), 64), device='cuda')
d_vs = torch.as_tensor(np.concatenate(ts)[:, :, None], device='cuda'); d_vp = torch.as_tensor(np.concatenate(ps), device='cuda'); d_vv = torch.as_tensor(np.concatenate(vs), device='cuda')
q = torch.as_tensor(np.ascontiguousarray(wl.cfg1_query_grid(1000).T), device='cuda')
sets = ((0, 'normal'), (4, 'no panel fill'), (1, 'no MMA'), (8, 'no reduce'), (13, 'none of them'))
if len(sys.argv) > 1: sets = tuple((int(a), f'flags {a}') for a in sys.argv[1:])
for flags, name in sets:
    lib.kmpx_debug_set(flags)
    for rep in range(3):
        kb.events = []
        kb.run_device(base_of, d_vs, d_vp, d_vv, q); torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for n, a, b in kb.events if n == 'kmpx_kmp_predict')
    print(f'{name:10s} predict {ms:.2f} ms for {len(base_of)} problems')
lib.kmpx_debug_set(0)
