"""Opcode sequence of an address range of one kernel: python tools/sass_seq.py obj kernel-substring lo hi"""
import re, subprocess, sys
obj, filt, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for blk in re.split(r"\n\s*Function : ", txt)[1:]:
    if filt not in blk.split("\n", 1)[0]:
        continue
    ops = []
    for l in blk.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and lo <= int(m.group(1), 16) <= hi:
            t = re.sub(r"^@!?U?P\d+\s+", "", m.group(2))
            ops.append(t.split()[0].split(".")[0] if not t.startswith("MUFU") else t.split()[0])
    for i in range(0, len(ops), 16):
        print(" ".join(ops[i:i + 16]))
    break
