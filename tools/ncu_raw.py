import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; units = rows[1]
want = ["Kernel Name","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers",
 "gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__warps_active.avg.pct_of_peak_sustained_active",
 "smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
 "smsp__thread_inst_executed_per_inst_executed.ratio","sm__cycles_elapsed.max","launch__shared_mem_per_block_dynamic","smsp__inst_executed_op_shared_ld.sum","smsp__inst_executed_op_shared_st.sum",
 "sm__sass_thread_inst_executed_op_dfma_pred_on.sum","smsp__sass_thread_inst_executed_op_dfma_pred_on.sum","local_load","local_store"]
for r in rows[2:]:
    print("-----")
    for h,u,v in zip(hdr,units,r):
        if h in want or "stalled" in h and "per_issue_active" in h and "not_issued" not in h or "local" in h and "sum" in h and "inst" in h:
            print(f"  {h} = {v} {u}")
