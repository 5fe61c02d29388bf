"""Operand-bandwidth model of a kernel's main loop from SASS: FP64 instructions with 3 distinct non-reused
register source pairs issue every 3 cycles on sm_100a, all others every 2 (profiles/r01_micro_fp64_pipe.txt)."""
import re, subprocess, sys, collections
so = "mpconstellation_b200/csrc/libmpc_b200.so"
for pat in sys.argv[1:]:
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    f = [x for x in re.split(r"\n\s*Function : ", txt)[1:] if pat in x.split("\n", 1)[0]][0]
    ins = []
    for l in f.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2)))
    back = []
    for a, t in ins:
        m = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a: back.append((a, int(m.group(1), 16)))
    a_end, a_start = max(back, key=lambda p: p[0] - p[1])
    n3 = n2 = other = 0
    for a, t in ins:
        if not (a_start <= a <= a_end): continue
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0]
        if op.split(".")[0] in ("DFMA", "DMUL", "DADD"):
            srcs = t[len(op):].split(",")[1:]
            regs = set(re.findall(r"\bR(\d+)", ",".join(o for o in srcs if ".reuse" not in o)))
            if len(regs) >= 3: n3 += 1
            else: n2 += 1
        else: other += 1
    print(f"{pat}: loop FP64 {n3+n2} (3-distinct-operand {n3}, <=2 {n2}), other {other}; operand-limited floor {3*n3+2*n2} cycles/warp-iteration (all-2-cycle floor {2*(n3+n2)})")
