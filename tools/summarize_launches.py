"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel / grid / block.

    python tools/summarize_launches.py gpurun_out/r02_launches.csv "header comment" > profiles/r02_launches_summary.csv
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, r["Grid Size"], r["Block Size"], float(r["Metric Value"].replace(",", "")) / 1e6))
    agg = defaultdict(lambda: [0, 0.0])
    for name, grid, block, ms in rows:
        a = agg[(name, grid, block)]
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    if note:
        print(f"# {note}")
    print(f"# {len(rows)} launches, {total:.3f} ms summed device time (cold-cache, serialised under ncu: compare shares)")
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "grid", "block", "launches", "total_ms", "share"])
    for (name, grid, block), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([name, grid, block, n, f"{ms:.3f}", f"{ms / total:.4f}"])


if __name__ == "__main__":
    main()
