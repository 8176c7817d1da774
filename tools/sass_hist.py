"""SASS loop finder + opcode histogram of one kernel in an object file: python tools/sass_hist.py obj kernel-substring"""
import collections, re, subprocess, sys
obj, filt = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if filt not in name:
        continue
    ins = []
    for l in blk.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print("==", name, len(ins), "instructions")
    loops = []
    for a, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and (a - int(m.group(1), 16)) // 16 > 20 and not t.startswith("BRA"):
                loops.append((int(m.group(1), 16), a))
    def hist(lo, hi):
        c = collections.Counter()
        for a, t in ins:
            if lo <= a <= hi:
                t = re.sub(r"^@!?U?P\d+\s+", "", t)
                op = t.split()[0]
                op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "SHFL", "MUFU", "RED", "LDG", "STG", "LDL", "STL")) else op.split(".")[0]
                c[op] += 1
        return c
    for lo, hi in loops:
        c = hist(lo, hi)
        print("loop %#x..%#x: %d instr" % (lo, hi, sum(c.values())))
        print("   ", ", ".join("%s %d" % kv for kv in c.most_common()))
