"""Summarise an .ncu-rep (first kernel row) into the handful of metrics we track.  usage: ncu_summary.py rep [out.json]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
out = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    def f(k):
        try: return float(d[k].replace(",", ""))
        except Exception: return None
    s = {
        "kernel": d.get("Kernel Name"), "grid": d.get("launch__grid_size"), "regs": f("launch__registers_per_thread"),
        # ncu reports gpu__time_duration.sum in the unit of the CSV's second header row (us for these kernels); keep
        # the number with its unit instead of guessing (round-1 summaries labelled microseconds as "duration_ms")
        "duration": f("gpu__time_duration.sum"),
        "duration_unit": units[hdr.index("gpu__time_duration.sum")] if "gpu__time_duration.sum" in hdr else None,
        "dram_read_MB": f("dram__bytes_read.sum"), "dram_write_MB": f("dram__bytes_write.sum"),
        "dram_pct_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_throughput_pct": f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "inst_executed": f("smsp__inst_executed.sum"),
        "thread_inst_per_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "smem_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "smem_bank_conflicts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "smem_bank_conflicts_st": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
        "smem_bank_conflicts_ld": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
        "local_ld_requests": f("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum"),
        "stalls_pcsamp": {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): f(k) for k in hdr
                          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and (f(k) or 0) > 0},
    }
    if s["dram_read_MB"] is not None:
        ur, uw = d and units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        s["dram_units"] = [ur, uw]
    out.append(s)
txt = json.dumps(out, indent=1)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
