"""Summarise an `ncu --set full` capture for profiles/: the metrics DESIGN.md quotes, one per line.

    python scripts/ncu_summary.py gpurun_out/r2j_prof_parity.ncu-rep "header text" > profiles/r2j_parity_ncu_summary.txt
"""
import csv, io, subprocess, sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
        "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")

rep = sys.argv[1]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else rep))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
print("Kernel Name\t\t" + vals[hdr.index("Kernel Name")])
for h, u, v in sorted(zip(hdr, units, vals)):
    if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        print(f"{h}\t{u}\t{v}")
