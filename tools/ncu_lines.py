"""Per-source-line stall samples from an ncu report: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv;
python tools/ncu_lines.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, out, hdr = None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] not in ("", "Function Name") and r[2] == "-":
        try:
            out.append((int(r[6]), int(r[7]), cur_file, r[0], r[1].strip()))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
print("total samples", tot)
for s, n, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  inst={n:9d}  {f}:{ln:>4s}  {src[:110]}")
