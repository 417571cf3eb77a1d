#!/usr/bin/env python
"""Key counters of every kernel in an ncu report: ncu_summary.py X.ncu-rep [frames_per_launch]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum",
        "smsp__inst_executed.sum", "smsp__inst_executed_op_shfl.sum" , "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__icc_requests_lookup_miss.sum", "smsp__warps_issue_stalled_no_instruction.avg" ]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
for r in rows[2:]:
    print("##", r[hdr.index("Kernel Name")][:70])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            v = r[i]
            extra = ""
            if frames and w.endswith(".sum") and "bytes" not in w and "time" not in w:
                try:
                    extra = f"   ({float(v.replace(',', '')) / frames:.1f} / frame)"
                except ValueError:
                    pass
            print(f"  {w:75s} {v} {units[i]}{extra}")
    stalls = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]]
    tot = sum(s for s, _ in stalls) or 1
    print("  stalls:", ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {100 * s / tot:.1f}%" for s, h in sorted(stalls, reverse=True)[:9]))
