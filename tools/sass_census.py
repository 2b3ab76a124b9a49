"""Opcode census of every kernel in libkmpx.so (sm_100a SASS): FP64 tensor MMAs (DMMA), TMA loads (UTMALDG), scalar FP64
FMAs, shared / global loads, cp.async (LDGSTS), mbarrier operations (SYNCS), named/CTA barriers.  Evidence for which
kernels run on the tensor pipe and which are TMA-fed.  usage: python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'kmp_demos_b200', 'csrc', 'libkmpx.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ops = ['DMMA', 'UTMALDG', 'DFMA', 'DADD', 'DMUL', 'MUFU', 'LDS', 'STS', 'LDG', 'STG', 'LDGSTS', 'SYNCS', 'BAR', 'SHFL', 'PREFETCH', 'CCTL', 'LDL', 'STL']
cur, cnt, arch = None, collections.OrderedDict(), ''
for ln in out.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0].replace('void ', '').replace('kmpx::', '')
        cnt[cur] = collections.Counter(); continue
    m = re.search(r'arch = (sm_\w+)', ln)
    if m:
        arch = m.group(1)
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', ln)
    if m and cur:
        op = m.group(1)
        for o in ops:
            if op == o or (o in ('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL') and op.startswith(o) and not op.startswith('LDGSTS')) or (o == 'PREFETCH' and op == 'CCTL'):
                cnt[cur][o] += 1; break
        cnt[cur]['total'] += 1
print(f'# {os.path.basename(lib)}: {len(cnt)} kernels, {arch}; static SASS instruction counts per kernel (unrolled code, not executed counts)')
print(f'{"kernel":52s} ' + ' '.join(f'{o:>8s}' for o in ['total'] + ops[:-4] + ['LDL', 'STL']))
for k, c in cnt.items():
    print(f'{k[:52]:52s} ' + ' '.join(f'{c[o]:8d}' for o in ['total'] + ops[:-4] + ['LDL', 'STL']))
