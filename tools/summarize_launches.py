"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the captured time).
usage: python tools/summarize_launches.py launches.csv > summary.md"""
import collections
import csv
import re
import sys


def main(path):
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path, errors="ignore")):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] == "ns" else v * 1e3 if d["Metric Unit"] == "ms" else v
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        name = re.sub(r"^void ", "", name)[:100]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "clusten::" in k)
    print(f"captured launches: {sum(v[0] for v in agg.values())}, total device time {tot / 1e3:.2f} ms "
          f"(cold-cache, serialised under ncu: compare SHARES); libclusten_b200 share {100 * ours / tot:.1f} %\n")
    print("| us | share | launches | kernel |\n|---:|---:|---:|---|")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.002:
            continue
        print(f"| {v[1]:.1f} | {100 * v[1] / tot:.1f} % | {v[0]} | `{n}` |")


if __name__ == "__main__":
    main(sys.argv[1])
