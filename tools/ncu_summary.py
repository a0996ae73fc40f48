#!/usr/bin/env python
"""Summarise an .ncu-rep: one block per profiled launch with the metrics the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] [--stalls]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    stalls = "--stalls" in sys.argv
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pat and pat not in r[kn]:
            continue
        print("==", r[hdr.index("ID")], r[kn][:110])
        for k in KEYS:
            if k in hdr:
                print(f"   {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        if stalls:
            st = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i]]
            if not st:
                st = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if "warp_issue_stalled" in h and h.endswith(".pct") and r[i]]
            for v, h in sorted(st, reverse=True)[:8]:
                print(f"   stall {h:75s} {v:10.3f}")


if __name__ == "__main__":
    main()
