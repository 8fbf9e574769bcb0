"""profiles/ncu_traffic.json from an ncu launch list (CSV with dram__bytes_read.sum / dram__bytes_write.sum):
average DRAM bytes per launch for the kernel families bench.py reports rooflines for."""
import csv
import json
import sys
from collections import defaultdict

FAMILIES = {"igemm": ("igemm_kernel", "halo_conv_kernel"), "wgrad": ("wgrad_kernel", "halo_wgrad"),
            "adam": ("adam_kernel",), "td": ("td_epilogue_kernel",)}


def main(path, out):
    per = defaultdict(lambda: defaultdict(float))
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(r["Metric Value"].replace(",", ""))
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
            per[int(r["ID"])]["bytes"] += v * scale
            per[int(r["ID"])]["name"] = r["Kernel Name"]
    res = {}
    for fam, keys in FAMILIES.items():
        xs = [d["bytes"] for d in per.values() if any(k in d["name"] for k in keys)]
        if xs:
            res[fam] = sum(xs) / len(xs)
    json.dump(res, open(out, "w"))
    print(res)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
