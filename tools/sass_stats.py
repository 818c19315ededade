"""Static SASS statistics of one kernel of a built libuavca: instruction count up to the first unconditional EXIT
(the common path of the step kernels), per-opcode and per-source-line (needs -lineinfo).
    python tools/sass_stats.py LIB.so KERNEL_SUBSTRING [--lines]"""
import re, subprocess, sys, tempfile, os
from collections import Counter
lib, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.startswith("uavca_kernels.")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
inside = False; cur = None; n = 0; ops = Counter(); cnt = Counter(); listing = []
for l in txt:
    if l.startswith(".text."):
        inside = pat in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m:
        n += 1
        op = m.group(2).split()
        pred = op[0].startswith("@")
        if pred: op = op[1:]
        ops[op[0].split(".")[0]] += 1
        cnt[cur] += 1
        listing.append((cur, m.group(2)))
        if op[0] == "EXIT" and not pred: break
print("instructions to first EXIT:", n)
print(" ".join(f"{k}:{v}" for k, v in ops.most_common()))
if "--lines" in sys.argv:
    for k, v in sorted(cnt.items(), key=lambda kv: (kv[0] or ("", 0))): print(k, v)
if "--list" in sys.argv:
    for c, t in listing: print(f"{c[0]}:{c[1]:<5d} {t}" if c else t)
