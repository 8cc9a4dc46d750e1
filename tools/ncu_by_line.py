"""Joins an `ncu --page source --csv` SASS listing with nvdisasm line info and
aggregates instructions executed / stall samples per CUDA source line.
usage: python tools/ncu_by_line.py report.ncu-rep cubin mangled_kernel_substring [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.check_output(["nvdisasm", "-g", "-c", cubin]).decode(errors="ignore").splitlines()
# walk the listing of the wanted function: remember the current line annotation
lines_for_instr = []
infn = False
cur = None
for ln in dis:
    if ln.startswith(".text.") or ".section" in ln and ".text." in ln:
        infn = kern in ln
    if not infn:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines_for_instr.append(cur)
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"]).decode(errors="ignore")
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
print("sass instrs: ncu %d, nvdisasm %d" % (len(body), len(lines_for_instr)))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = tot_samp = 0
for i, r in enumerate(body):
    key = lines_for_instr[i] if i < len(lines_for_instr) else None
    inst = int(r[col["Instructions Executed"]]); samp = int(r[col["# Samples"]])
    agg[key][0] += inst; agg[key][1] += samp
    tot_inst += inst; tot_samp += samp
    for s in stallcols:
        v = int(r[col[s]])
        if v:
            agg[key][2][s] += v
print("total warp instructions %d, samples %d" % (tot_inst, tot_samp))
for key, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    tops = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%-28s inst %5.1f%%  samples %5.1f%%   %s" % (key, 100. * inst / tot_inst, 100. * samp / tot_samp, tops))
