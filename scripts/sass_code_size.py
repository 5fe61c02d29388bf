"""Code size of a kernel attributed to its source lines: instructions per (file, 10-line bucket) from
`nvdisasm --print-line-info` of the built library (compiled with -lineinfo).  Used for the instruction-cache analysis of
the adaptive kernel.   python scripts/sass_code_size.py <substring of the mangled kernel name>"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mpconstellation_b200", "csrc", "libmpc_b200.so")
want = sys.argv[1] if len(sys.argv) > 1 else "discretize_adaptive_kernelILb0ELi32ELi1ELb0ELb0"
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cur, on, cnt = None, False, collections.Counter()
for line in text.splitlines():
    if line.startswith(".text."):
        on = want in line
        if on:
            print("kernel:", line[6:-1])
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)) // 10 * 10)
    elif re.search(r"/\*[0-9a-f]{4,5}\*/", line):
        cnt[cur] += 1
tot = sum(cnt.values())
print(f"{tot} instructions = {tot * 16 / 1024:.0f} KiB")
for (f, l), c in sorted(cnt.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if c >= 0.005 * tot:
        print(f"{c:6d}  {100.0 * c / tot:5.1f} %  {f}:{l}-{l + 9}")
