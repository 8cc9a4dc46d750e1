"""SASS of one kernel of a built object, joined with its line info (nvdisasm -g): instruction mix,
spills and branches per source line.  Runs in the build container (no GPU needed).

usage: python tools/sass_report.py quantpde_b200/lib/obj/qpde_lean.o k_lean_sweep [--file qpde_lean.cu --lines 300:480] [--dump]
"""
import argparse
import collections
import os
import re
import subprocess
import tempfile


def kernel_sass(obj, kernel):
    """[(file, line, opcode, text)] of every instruction of the first kernel whose name contains `kernel`."""
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
        out = []
        for cubin in sorted(os.listdir(tmp)):
            txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
            inside, cur, name = False, None, None
            for ln in txt.splitlines():
                m = re.match(r"\s*\.text\.(\S+):", ln)
                if m:
                    if inside and out:
                        return name, out
                    inside = kernel in m.group(1); name = m.group(1)
                    continue
                if not inside:
                    continue
                m = re.search(r'//## File "([^"]*)", line (\d+)', ln)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
                m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)(.*?);", ln)
                if m and cur:
                    out.append((cur[0], cur[1], m.group(2), ln.strip()))
            if inside and out:
                return name, out
    raise SystemExit("kernel %s not found in %s" % (kernel, obj))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj"); ap.add_argument("kernel")
    ap.add_argument("--file"); ap.add_argument("--lines")
    ap.add_argument("--dump", action="store_true")
    a = ap.parse_args()
    name, ins = kernel_sass(a.obj, a.kernel)
    print("kernel", name, "instructions", len(ins))
    lo, hi = (int(v) for v in a.lines.split(":")) if a.lines else (0, 1 << 30)
    sel = [i for i in ins if (not a.file or i[0] == a.file) and lo <= i[1] <= hi]
    mix = collections.Counter(i[2].split(".")[0] for i in sel)
    print("selected", len(sel), " ".join("%s %d" % kv for kv in mix.most_common(24)))
    spills = collections.Counter((i[0], i[1]) for i in ins if i[2].startswith(("STL", "LDL")))
    print("spill instructions by line:", dict(spills))
    if a.dump:
        per = collections.OrderedDict()
        for f, l, op, _ in sel:
            per.setdefault((f, l), collections.Counter())[op.split(".")[0]] += 1
        for (f, l), c in sorted(per.items()):
            print("%s:%d  %d  %s" % (f, l, sum(c.values()), dict(c.most_common(6))))


if __name__ == "__main__":
    main()
