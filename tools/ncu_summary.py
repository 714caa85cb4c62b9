"""Summarise an .ncu-rep: key metrics + stall breakdown (run on the CPU box: `python tools/ncu_summary.py rep`)."""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    g = lambda k: d.get(k, ("nan", ""))[0]
    print("kernel:", g("Kernel Name")[:90])
    for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__occupancy_limit_registers",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__sass_inst_executed_op_shared_ld.sum"]:
        if k in d: print(f"  {k:75s} {d[k][0]:>18s} {d[k][1]}")
    st = {k: float(v[0].replace(",", "")) for k, v in d.items() if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k)}
    tot = sum(st.values())
    print("  stall reasons (warps per issue-active cycle), total %.2f:" % tot)
    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:12]:
        print(f"    {k.split('stalled_')[1].split('_per_issue')[0]:28s} {v:6.3f}  {100*v/tot:5.1f}%")
