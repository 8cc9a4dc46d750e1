"""Per CUDA-source-line opcode breakdown of an ncu report (instructions executed
per warp per contract / per solve).  usage: ncu_ops_by_line.py rep cubin kernel contracts warps solves [top]"""
import csv, re, subprocess, sys
from collections import defaultdict
rep, cubin, kern = sys.argv[1:4]
contracts, warps, solves = int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6])
top = int(sys.argv[7]) if len(sys.argv) > 7 else 40
dis = subprocess.check_output(["nvdisasm", "-g", "-c", cubin]).decode(errors="ignore").splitlines()
lines = []; infn = False; cur = None
for ln in dis:
    if ".section" in ln and ".text." in ln: infn = kern in ln
    if not infn: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append(cur)
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"]).decode(errors="ignore")
rows = list(csv.reader(src.splitlines())); hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: defaultdict(int)); tot = defaultdict(int); samp = defaultdict(int); ops = defaultdict(int)
for i, r in enumerate(rows[2:]):
    ins = r[col['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ins)
    op = m.group(2).split('.')[0] if m else ins
    n = int(r[col['Instructions Executed']])
    agg[lines[i]][op] += n; tot[lines[i]] += n; ops[op] += n; samp[lines[i]] += int(r[col['# Samples']])
T = sum(tot.values()); S = sum(samp.values()); norm = contracts * warps * solves
print("instructions per warp per solve: %.0f" % (T / norm))
print(' '.join('%s %.0f' % (k, v / norm) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:28]))
text = open('quantpde_b200/csrc/qpde_resident.cu').read().splitlines()
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    o = ' '.join('%s:%.0f' % (a, c / norm) for a, c in sorted(agg[k].items(), key=lambda kv: -kv[1])[:6])
    t = text[k[1] - 1].strip()[:58] if k and k[0] == 'qpde_resident.cu' else ''
    print('%-28s inst %4.1f%% %4.0f smp %4.1f%% | %s | %s' % (k, 100 * v / T, v / norm, 100 * samp[k] / S, o, t))
