"""Per-CUDA-source-line stall-sample summary of one kernel from an .ncu-rep (needs -lineinfo builds and
ncu --import-source on).  usage: python tools/ncu_hot.py <report> <kernel regex> [top]"""
import csv, subprocess, sys, io, os

def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                          f'regex:{kern}'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, hdr, items = '', None, []
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            fpath = os.path.basename(r[1]); continue
        if len(r) > 6 and r[0] == 'Line No':
            hdr = r; si = hdr.index('# Samples'); continue
        if hdr is None or len(r) <= si:
            continue
        if r[0].strip().isdigit() and r[2].strip() == '-':
            s = r[si].strip()
            if s.isdigit():
                items.append((int(s), fpath, r[0].strip(), r[1].strip()))
    tot = sum(i[0] for i in items) or 1
    print('total samples', tot)
    for s, f, ln, text in sorted(items, reverse=True)[:top]:
        print(f'{s:8d} {s / tot:6.1%}  {f}:{ln:<5s} {text[:110]}')

main()
