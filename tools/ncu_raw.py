"""Print selected raw metrics of every kernel in an .ncu-rep (read here, no GPU needed).
usage: python tools/ncu_raw.py <rep> <metric-substring>..."""
import csv
import subprocess
import sys


def main():
    rep, pats = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "")[:70], d.get("Grid Size"), d.get("Block Size"))
        for i, h in enumerate(hdr):
            if any(w in h for w in pats):
                print("   ", h, "=", r[i], units[i])


if __name__ == "__main__":
    main()
