#!/usr/bin/env python
"""One line per profiled kernel from an .ncu-rep (ncu --page raw --csv)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rdGB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("smsp__inst_executed.sum", "inst"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "st_sect"), ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "st_req"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sect"), ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_req")]
print(" | ".join(c[1] for c in cols))
for r in rows[2:]:
    out = []
    for k, n in cols:
        v = r[idx[k]] if k in idx else "-"
        if n == "kernel": v = v.split("(")[0][-34:]
        else:
            try: v = "%.4g" % float(v.replace(",", ""))
            except Exception: pass
        out.append(v)
    print(" | ".join(out))
