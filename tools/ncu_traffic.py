"""Aggregate an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list
into profiles/traffic.json: DRAM bytes per launch of the dominant kernels (bench.py reads it for roofline.traffic).
usage: python tools/ncu_traffic.py <unet_traffic.csv> [<mix_traffic.csv>] <tag>"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(dict)          # launch id -> metric -> value
    names = {}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]].lower()
        m = r[idx["Metric Name"]]
        if "byte" in u:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        elif u.startswith(("ns", "nsecond")):
            v *= 1e-3
        elif u.startswith(("ms", "msecond")):
            v *= 1e3
        per[r[idx["ID"]]][m] = v
        names[r[idx["ID"]]] = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    return per, names


def family(per, names, pat):
    ids = [i for i, n in names.items() if re.search(pat, n)]
    if not ids:
        return None
    rd = sum(per[i].get("dram__bytes_read.sum", 0.0) for i in ids)
    wr = sum(per[i].get("dram__bytes_write.sum", 0.0) for i in ids)
    us = sum(per[i].get("gpu__time_duration.sum", 0.0) for i in ids)
    return dict(launches=len(ids), dram_read_bytes=rd, dram_write_bytes=wr, dram_bytes_per_launch=(rd + wr) / len(ids),
                total_us_under_ncu=us)


def main():
    args = sys.argv[1:]
    tag = args[-1]
    per, names = load(args[0])
    out = {}
    for key, pat in (("gemm", r"gemm_tc_kernel"), ("attention", r"attn_"), ("groupnorm", r"gn_"), ("layernorm", r"ln_kernel")):
        f = family(per, names, pat)
        if f:
            f["source"] = f"ncu dram__bytes_read+write over one CFG-batch-2 UNet forward @128x128 ({os.path.basename(args[0])}, {tag})"
            out[key] = f
    if len(args) > 2:
        per, names = load(args[1])
        f = family(per, names, r"slerp")
        if f:
            f["source"] = f"ncu dram__bytes_read+write, batched mix 2048 x 65536 fp16 ({os.path.basename(args[1])}, {tag})"
            out["mix"] = f
    fp = os.path.join(ROOT, "profiles", "traffic.json")
    with open(fp, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
