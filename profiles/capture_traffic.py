"""Writes profiles/traffic_r2.json: DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant
kernels from committed-build `ncu --set full` captures, with the hash of the kernel sources they were taken on (bench.py
compares it with the current build).   usage (CPU box, after the captures came back in gpurun_out/):
    python profiles/capture_traffic.py gemm_f16x3=gpurun_out/gemm.ncu-rep avg_heads=gpurun_out/avg.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def first_kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    val = lambda k: float(d[k].replace(",", "")) * UNIT[u[k]]
    return {"kernel": d["Kernel Name"][:120], "dram_bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
            "dram_read_bytes": int(val("dram__bytes_read.sum")), "dram_write_bytes": int(val("dram__bytes_write.sum")),
            "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(u["gpu__time_duration.sum"], 1),
            "grid": d.get("launch__grid_size"), "source": "profiles/" + os.path.basename(path)}


def main(args):
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    out = {"build": bench.csrc_hash()}
    for a in args:
        key, path = a.split("=")
        out[key] = first_kernel(path)
        out[key]["sources"] = list(bench.KERNEL_SOURCES.get(key, ()))
        out[key]["sources_hash"] = bench.kernel_sources_hash(key)
    extra = os.environ.get("TRAFFIC_NOTE")
    if extra:
        out["note"] = extra
    with open(os.path.join(ROOT, "profiles", "traffic_r2.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
