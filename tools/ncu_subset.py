"""Writes the subset of an ncu report's raw page that profiles/ keeps (metric,unit,value of the first
captured launch).  usage: python tools/ncu_subset.py report.ncu-rep out.csv"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes(_read|_write)?\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?$|dram__cycles_active\.avg"
    r"|gpu__dram_throughput\.avg\.pct|gpu__time_duration\.sum"
    r"|launch__(grid_size|block_size|registers_per_thread$|shared_mem_per_block_dynamic|occupancy_limit|waves|cluster_size)"
    r"|lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|lts__t_sectors_srcunit_tex_op_(read|write)\.sum$"
    r"|sm__(throughput\.avg\.pct|warps_active\.avg\.pct|cycles_elapsed\.max|pipe_fp64_cycles_active\.avg|inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active)"
    r"|smsp__(inst_executed\.sum$|issue_active\.avg\.pct|pcsamp_warps_issue_stalled)"
    r"|l1tex__data_pipe_lsu_wavefronts_mem_shared(\.sum$|_op_ld\.sum$|_op_st\.sum$)|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$"
    r"|smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on\.sum\.per_cycle_elapsed)")

rep, out = sys.argv[1:3]
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode(errors="ignore")
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
with open(out, "w") as f:
    f.write("metric,unit,value\n")
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            f.write("kernel,,\"%s\"\n" % v)
        if KEEP.match(h) and "not_issued" not in h:
            f.write("%s,%s,%s\n" % (h, u, v))
print(open(out).read())
