#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): key throughput / stall metrics per kernel, as text."""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct", "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum ", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled", "sm__warps_active.avg.pct", "sm__throughput.avg.pct", "lts__t_bytes.sum ", "lts__t_sectors_srcunit_tex_op_read.sum "]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("==", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for h in hdr:
        hh = h + " "
        if any(k in hh for k in KEYS):
            v = d[h]
            if "issue_stalled" in h and "ratio" in h:
                try:
                    if float(v) < 0.05: continue
                except ValueError: pass
            print(f"  {h} = {v} {u[h]}")
