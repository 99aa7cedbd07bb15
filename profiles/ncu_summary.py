"""Compact markdown summary of an `ncu --set full` report (read here, on the CPU box, with `ncu -i`).
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/x_ncu.md"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared mem)"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2 -> SM read rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe (of elapsed cycles)"),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "legacy HMMA pipe (of active cycles)"),
    ("SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "XU pipe (conversions)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
]


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print(f"## {path}\n\n(no kernels in the report)\n")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            print(f"## `{d.get('Kernel Name', '?')[:110]}`  ({path.split('/')[-1]})\n")
            print("| metric | value |\n|---|---|")
            for k, label in KEYS:
                if k in d and d[k] != "":
                    print(f"| {label} (`{k}`) | {d[k]} {u.get(k, '')} |")
            print()


if __name__ == "__main__":
    main(sys.argv[1:])
