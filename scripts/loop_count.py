"""Static size of the dequant loop: SASS instruction histogram of the innermost backward-branch loops that contain the
tcgen05.st (STTM) pair of tc_gemm_kernel<16,false,false>.  usage: python scripts/loop_count.py [path/to/tc_gemm.o]"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ferrum-infer-rs_b200", "csrc", "build", "tc_gemm.o")
elf = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
sym = sorted(set(re.findall(r"_ZN3fb2\w*tc_gemm_kernelILi16ELb0ELb0EEEvNS0_8TcParamsE14CUtensorMap_st\b", elf)))[0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, obj], capture_output=True, text=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
print(obj, "-- kernel instructions:", len(ins))
idx = {a: i for i, (a, _) in enumerate(ins)}
sttm = [i for i, (_, t) in enumerate(ins) if "STTM" in t]
loops = []
for i, (_, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) in idx and idx[int(m.group(1), 16)] < i:
        loops.append((idx[int(m.group(1), 16)], i))
for b, e in sorted(loops, key=lambda be: be[1] - be[0]):
    inside = [s for s in sttm if b <= s <= e]
    if inside and (e - b) < 1000:          # the dequant loop itself, not the segment / tile loops around it
        ops = Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for _, t in ins[b:e + 1])
        print(f"loop [{b},{e}]: {e - b + 1} instructions, {len(inside)} STTM ({len(inside) // 2} units):", dict(ops.most_common(16)))
