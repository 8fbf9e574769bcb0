"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name,
share of the captured window, and (optionally) every launch in order."""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((int(r["ID"]), name, v * scale, r.get("Grid Size", ""), r.get("Block Size", "")))
    return rows


def main():
    rows = load(sys.argv[1])
    tot = sum(r[2] for r in rows)
    agg = OrderedDict()
    for _id, name, us, *_ in rows:
        t, n = agg.get(name, (0.0, 0))
        agg[name] = (t + us, n + 1)
    print(f"{len(rows)} launches, {tot/1e3:.3f} ms total")
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={n:4d}  avg {t/n:8.1f} us  {name[:90]}")
    if len(sys.argv) > 2:
        for _id, name, us, grid, blk in rows:
            print(f"{_id:6d} {us:9.1f} us grid={grid:>14s} {name[:80]}")


if __name__ == "__main__":
    main()
