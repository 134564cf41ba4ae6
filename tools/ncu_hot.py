"""Top stall locations of an `ncu --page source --csv --print-source sass` export.  usage: python tools/ncu_hot.py src.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.2
hdr = rows[1]
iS, iSamp, iEx = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
seen = {}
for i in stall_cols:
    seen.setdefault(hdr[i], i)       # first occurrence = all samples
data = []
for r in rows[2:]:
    try:
        data.append((int(r[iSamp]), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, (s, r) in enumerate(data):
    if s > tot * minp / 100:
        st = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for h, i in seen.items()), reverse=True)[:2]
        print(f"{n:4d} {s:6d} {100*s/tot:5.1f}%  ex={r[iEx]:>8s} {r[iS].strip()[:64]:64s} {st}")
