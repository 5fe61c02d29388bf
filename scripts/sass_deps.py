"""Static ILP estimate of a kernel loop from SASS: for every FP64 instruction, distance (in FP64-pipe
instructions) to the closest earlier instruction that wrote one of its source registers."""
import re, subprocess, sys, collections
so = "mpconstellation_b200/csrc/libmpc_b200.so"; pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
f = [x for x in re.split(r"\n\s*Function : ", txt)[1:] if pat in x.split("\n", 1)[0]][0]
ins = []
for l in f.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
back = [(a, int(re.search(r"0x([0-9a-f]+)", t).group(1), 16)) for a, t in ins if t.split()[0].startswith("BRA") or " BRA " in t if re.search(r"0x([0-9a-f]+)", t)]
back = [(a, tgt) for a, tgt in back if tgt < a]
a_end, a_start = max(back, key=lambda p: p[0] - p[1])
body = [(a, t) for a, t in ins if a_start <= a <= a_end]
def regs(tok):
    out = []
    for r in re.findall(r"R(\d+)", tok):
        out.append(int(r))
    return out
last_write = {}; fp_idx = 0; dist = collections.Counter(); cyc_model = 0
lat_hist = []
for a, t in body:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0].split(".")[0]
    ops = t[len(t.split()[0]):].split(",")
    is_fp = op in ("DFMA", "DMUL", "DADD")
    if op in ("DFMA", "DMUL", "DADD", "MUFU", "LDS", "MOV", "FSEL", "SEL", "IMAD", "DSETP"):
        dst = regs(ops[0]); srcs = [r for o in ops[1:] for r in regs(o)]
        # 64-bit: register pairs
        srcp = set()
        for r in srcs: srcp.add(r); srcp.add(r + 1)
        if is_fp:
            d = min([fp_idx - last_write[r][0] for r in srcp if r in last_write and last_write[r][1]] or [99])
            dist[min(d, 12)] += 1
        for r in dst:
            last_write[r] = (fp_idx, is_fp); last_write[r + 1] = (fp_idx, is_fp)
    if is_fp: fp_idx += 1
tot = sum(dist.values())
print(pat, "FP64 instr in loop:", tot)
for d in sorted(dist): print(f"  producer distance {d:2d}{'+' if d==12 else ' '}: {dist[d]:4d} ({100*dist[d]/tot:4.1f}%)")
