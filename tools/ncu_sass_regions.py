"""Group the SASS of an `ncu --page source --csv` export into regions of equal execution count (loop bodies) and print, per
region, instructions executed (total and share), stall samples and an opcode histogram.   usage: python tools/ncu_sass_regions.py src.csv [min_share]"""
import collections
import csv
import sys


def main(path, min_share=0.01):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hi]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows[hi + 1:] if len(r) > iex and r[iex].isdigit()]
    total = sum(int(r[iex]) for r in body)
    tot_s = sum(int(r[ismp]) for r in body)
    print(f"SASS instructions {len(body)}, executed {total}, samples {tot_s}")
    regions, cur = [], None
    for n, r in enumerate(body):
        ex = int(r[iex])
        if cur is None or abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"], 1):
            cur = dict(start=n, ex=ex, n=0, sum=0, smp=0, ops=collections.Counter())
            regions.append(cur)
        cur["n"] += 1; cur["sum"] += ex; cur["smp"] += int(r[ismp])
        cur["ops"][r[isrc].split()[0 if not r[isrc].strip().startswith("@") else 1].split(".")[0]] += 1
    for g in regions:
        if g["sum"] >= min_share * total:
            ops = " ".join(f"{k}:{v}" for k, v in g["ops"].most_common(12))
            print(f"[{g['start']:5d}+{g['n']:4d}] exec/instr {g['ex']:>10d}  share {100 * g['sum'] / total:5.1f}%  samples {100 * g['smp'] / max(tot_s, 1):5.1f}%  {ops}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.01)
