"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (cold-cache, serialised times:
compare SHARES, not absolutes).  usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/x.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")
    print(f"\ntotal {tot / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
