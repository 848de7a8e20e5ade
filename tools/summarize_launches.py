"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name:
   python tools/summarize_launches.py gpurun_out/train_launches.csv [top_n] > profiles/<name>.txt"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    ki, vi, ui = header.index("Kernel Name"), header.index("Metric Value"), header.index("Metric Unit")
    mi = header.index("Metric Name")
    for r in rd:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        ns = v * {"nsecond": 1, "ns": 1, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9}.get(unit, 1)
        rows.append((r[ki], ns))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        k = re.sub(r"\(.*$", "", k)
        agg[k][0] += 1
        agg[k][1] += ns
    total = sum(v[1] for v in agg.values())
    print("launches %d, kernels %d, total device time %.3f ms (cold-cache, serialised: compare shares)" % (len(rows), len(agg), total / 1e6))
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%6.2f%% %9.3f ms %5d x  %s" % (100 * ns / total, ns / 1e6, n, k[:150]))


if __name__ == "__main__":
    main()
