"""Table-driven predict kernel (predict_tab.cu, default) vs the cursor-driven one (predict_fast.cu, kmpx_debug_set(4096)): agreement and time.
usage: tab_probe.py [sets per letter] [debug flag sets ...]"""
import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmp_demos_b200 import _lib, workloads as wl
from kmp_demos_b200.batch import KMPBatch
lib = _lib.load()
mu, sig, _ = wl.load_letter_references()
L = 26; c = wl.CFG1
NB = int(os.environ.get('TAB_PROBE_N', '200'))          # stored points of the base trajectories
nset = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flagsets = [int(a) for a in sys.argv[2:]] or [4096, 0]
ts, ps, vs = [], [], []
for li in range(L):
    t, p, v = wl.cfg3_via_sets(li, mu[li], nset); ts.append(t); ps.append(p); vs.append(v)
base_of = torch.as_tensor(np.repeat(np.arange(L, dtype=np.int32), nset), device='cuda')
d_vs = torch.as_tensor(np.concatenate(ts)[:, :, None], device='cuda'); d_vp = torch.as_tensor(np.concatenate(ps), device='cuda'); d_vv = torch.as_tensor(np.concatenate(vs), device='cuda')
q = torch.as_tensor(np.ascontiguousarray(wl.cfg1_query_grid(1000).T), device='cuda')
out = {}
kb = KMPBatch(l=c['lam'], alpha=c['alpha'], sigma_f=c['sigma_f'], chunk=5328)
kb.set_base(np.broadcast_to(wl.cfg1_kmp_grid(NB), (L, 1, NB)), np.ascontiguousarray(mu[:, :, :NB]), np.ascontiguousarray(sig[..., :NB]))
for fl in flagsets:
    lib.kmpx_predict_tile_queries(32 if fl & 262144 else (64 if fl & 524288 else 0))     # + 262144: 32-query CTAs, + 524288: 64-query CTAs
    lib.kmpx_debug_set(fl & ~(262144 | 524288))
    for rep in range(3):
        kb.events = []
        xi, sg, info, _ = kb.run_device(base_of, d_vs, d_vp, d_vv, q); torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for n, a, b in kb.events if n.startswith('kmpx_kmp_predict'))
    st = (ctypes.c_longlong * 2)(); lib.kmpx_debug_predict_stats(st, 1)
    print(f'flags {fl:5d} predict {ms:.3f} ms for {len(base_of)} problems ({ms / len(base_of) * 26624:.1f} ms per 26624); k-steps executed / scheduled {st[0] / max(st[1], 1):.3f}', flush=True)
    out[fl] = (xi.clone(), sg.clone())
lib.kmpx_debug_set(0); lib.kmpx_predict_tile_queries(0)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
for a in out:
    if a != flagsets[0]:
        print(f'flags {a} vs {flagsets[0]}: mean', rel(out[a][0], out[flagsets[0]][0]), 'cov', rel(out[a][1], out[flagsets[0]][1]))
