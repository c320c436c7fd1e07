#!/usr/bin/env python
"""Summarise one .ncu-rep (raw page CSV) into the handful of metrics that matter here.  Usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__cycles_active.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("-" * 100)
    for k in KEYS:
        for h in hdr:
            if h == k or h.endswith("." + k) or h.endswith(k):
                print(f"{h:90s} {d[h]:>16s} {units[hdr.index(h)]}")
                break
