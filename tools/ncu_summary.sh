#!/bin/bash
# usage: tools/ncu_summary.sh <rep.ncu-rep>   -> key counters (name, unit, value) from `--page raw --csv`
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units,vals=rows[0],rows[1],rows[2:]
want=["gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__shared_mem_per_block_dynamic","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","launch__occupancy_limit_warps","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fp64.sum","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.sum","sm__inst_executed_pipe_fma.sum","sm__inst_executed_pipe_xu.sum","sm__inst_executed_pipe_lsu.sum","sm__inst_executed_pipe_uniform.sum","sm__inst_executed_pipe_cbu.sum","sm__inst_executed_pipe_adu.sum","smsp__sass_thread_inst_executed_op_dfma_pred_on.sum","smsp__sass_thread_inst_executed_op_dmul_pred_on.sum","smsp__sass_thread_inst_executed_op_dadd_pred_on.sum","smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed","sm__cycles_elapsed.max","sass__inst_executed_local_loads","sass__inst_executed_local_stores","sass__inst_executed_shared_loads","sass__inst_executed_shared_stores","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","dram__bytes_read.sum","dram__bytes_write.sum"]
for k in range(len(vals)):
    print("== launch",k, vals[k][hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "")
    for w in want:
        if w in hdr:
            i=hdr.index(w); print(f"{w:85s} {vals[k][i]:>20s} {units[i]}")
    for i,h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and float(vals[k][i] or 0)>0.05:
            print(f"{h:85s} {vals[k][i]:>20s}")
'
