"""Text summary of an Nsight Compute report (dev tool): one block of headline metrics per captured launch.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--header "free text"] > profiles/rNN_xxx_ncu_full.txt
"""
import argparse
import csv
import io
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--header", action="append", default=[])
    ap.add_argument("--max", type=int, default=0, help="at most N launches per kernel name")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for h in a.header:
        print("# " + h)
    print(f"# source: {a.report} (ncu --set full --clock-control none --import-source on, B200, one GPU)")
    seen = {}
    for n, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]]
        seen[name] = seen.get(name, 0) + 1
        if a.max and seen[name] > a.max:
            continue
        print(f"\n[launch {n}] {name}")
        for m in METRICS:
            if m in idx:
                print(f"  {m:72s} {r[idx[m]]:>18s} {units[idx[m]]}")


if __name__ == "__main__":
    main()
