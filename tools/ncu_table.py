"""Compact table of an .ncu-rep (read here, no GPU needed): one row per profiled launch with the roofline-relevant
raw metrics.  usage: python tools/ncu_table.py <rep> [title] > profiles/<name>.md"""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("launch__registers_per_thread", "regs"),
        ("launch__waves_per_multiprocessor", "waves")]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n\nsource: `{rep}` (`ncu --set full --clock-control none`; per-launch, cold caches)\n")
    print("| kernel | grid | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].replace("<unnamed>::", "").split("(")[0]
        cells = []
        for m, _ in COLS:
            if m in idx and r[idx[m]] not in ("", "n/a"):
                cells.append(f"{r[idx[m]]} {units[idx[m]]}".strip())
            else:
                cells.append("-")
        print(f"| {name} | {r[idx['Grid Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
