"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
usage: python tools/ncu_summary.py gpurun_out/unet_launches.csv > profiles/rNN_unet_launches.md"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        val = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        val = val / 1000 if unit.startswith("n") else val * 1000 if unit.startswith("m") else val
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        key = (name, r[idx["Grid Size"]]) if "--by-grid" in sys.argv else (name,)
        agg[key][0] += 1
        agg[key][1] += val
    tot = sum(v[1] for v in agg.values())
    print(f"source: {path}\n")
    print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches "
          "(ncu per-launch times are cold-cache and serialised: compare SHARES)\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {' '.join(k)} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
