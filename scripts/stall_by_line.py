"""Attribute ncu stall samples (SASS level) to CUDA source lines using nvdisasm line info.
usage: stall_by_line.py rep.ncu-rep kernel-mangled-substring [so]     (STALL_LINES=n: print the n top lines, default 40; 0 = all)"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, pat = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else "mpconstellation_b200/csrc/libmpc_b200.so"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
sec = None; line = None; addr2line = {}
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m: sec = m.group(1); continue
    if sec is None or pat not in sec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*?);", l)
    if m: addr2line[int(m.group(1), 16)] = (line, m.group(2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = int(rd[0]["Address"], 16) if rd[0]["Address"].startswith("0x") else int(rd[0]["Address"])
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
for r in rd:
    a = (int(r["Address"], 16) if r["Address"].startswith("0x") else int(r["Address"])) - base
    ln = addr2line.get(a, (None, ""))[0]
    n = int(r["# Samples"] or 0); tot += n
    c = agg[ln]; c["samples"] += n; c["inst"] += int(r["Instructions Executed"] or 0)
    for k in ("stall_wait", "stall_math", "stall_short_sb", "stall_long_sb", "stall_not_selected", "stall_selected", "stall_dispatch"):
        c[k] += int(r.get(k) or 0)
srcfile = {}
def text(ln):
    if ln is None: return ""
    f, n = ln
    if f not in srcfile:
        p = os.path.join("mpconstellation_b200/csrc", f)
        srcfile[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return srcfile[f][n - 1].strip()[:80] if n - 1 < len(srcfile[f]) else ""
print(f"total samples {tot}")
for ln, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:(int(os.environ.get("STALL_LINES", "40")) or None)]:
    print(f"{100*c['samples']/tot:5.1f}% inst {c['inst']:>10d} wait {c['stall_wait']:6d} math {c['stall_math']:6d} ssb {c['stall_short_sb']:5d} sel {c['stall_selected']:6d} | {ln} {text(ln)}")
