"""Top SASS instructions by stall samples with their stall-reason split, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.   python tools/ncu_sass.py f.src.csv [top] [file-filter]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = sys.argv[3] if len(sys.argv) > 3 else None
hdr = None; cur_line = None; cur_file = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    if r[0] != '':
        cur_line = r[0]
    if r[2].startswith('0x'):
        d = dict(zip(hdr, r))
        try:
            n = int(d['# Samples'] or 0)
        except ValueError:
            continue
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v) > 0}
        out.append((n, cur_file, cur_line, r[3].strip()[:64], st, d['Instructions Executed']))
tot = sum(o[0] for o in out)
print('total samples', tot)
agg = {}
for n, f, l, sass, st, ie in out:
    for k, v in st.items():
        agg[k] = agg.get(k, 0) + v
print('by reason:', ' '.join(f'{k}={v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])))
sel = [o for o in out if flt is None or flt in o[1]]
for n, f, l, sass, st, ie in sorted(sel, key=lambda o: -o[0])[:top]:
    s = ' '.join(f'{k}={v}' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{n:5d} {f[:14]:14s}:{l:>4s} ie={ie:>8s} {sass:64s} | {s}')
