"""Key counters of an ncu raw-page CSV: python tools/ncu_raw.py gpurun_out/X.raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "")[:60])
    for k in KEYS:
        if k in d:
            print(f"  {k:70s} {d[k]}")
    for k in hdr:
        if "warps_issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            v = float(d[k] or 0)
            if v > 0.15:
                print(f"  stall {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio',''):30s} {v:.2f}")
