"""Per-kernel SASS statistics of libmpc_b200.so: instruction mix of every loop body (cuobjdump -sass)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "mpconstellation_b200/csrc/libmpc_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "discretize_kernelILb0ELi128ELi1E"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for l in f.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name, len(ins), "instructions")

    def hist(lo, hi):
        c = collections.Counter()
        for a, t in ins:
            if lo <= a <= hi:
                t = re.sub(r"^@!?U?P\d+\s+", "", t)
                c[t.split()[0].split(".")[0]] += 1
        return c
    for a, t in ins:
        m = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            c = hist(lo, a)
            tot = sum(c.values())
            fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
            print(f"  loop {hex(lo)}..{hex(a)}: {tot} instr, FP64 {fp64} ({100*fp64/tot:.0f}%)")
            print("   ", dict(c.most_common(16)))
