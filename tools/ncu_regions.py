"""Per-region (by execution count) instruction / stall-sample shares of one kernel in an .ncu-rep:
    python tools/ncu_regions.py rep.ncu-rep kernel-substring [top]"""
import collections, csv, io, re, subprocess, sys
rep, filt = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name" and filt in rows[i][1]:
        break
    i += 1
hdr = rows[i + 1]
data = []
stallcols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[i + 2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        data.append((int(d["Address"], 16), d["Source"].strip(), int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0),
                     {hdr[k][6:]: int(r[k] or 0) for k in stallcols if int(r[k] or 0) > 0}))
    except ValueError:
        pass
base = data[0][0]
tot, toti = sum(x[2] for x in data), sum(x[3] for x in data)
print("samples", tot, "warp-instructions", toti)
reg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), collections.Counter()])
for a, s, sm, ie, st in data:
    r = reg[ie]
    r[0] += 1; r[1] += sm; r[2] += ie
    for k, v in st.items():
        r[3][k] += v
    op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "SHFL", "MUFU", "RED", "LDG", "STG")) else op.split(".")[0]
    r[4][op] += 1
for ie, r in sorted(reg.items(), key=lambda kv: -kv[1][2])[:6]:
    st = sum(r[3].values()) or 1
    print("exec %9d  n_instr %5d  instr %5.1f%%  samples %5.1f%%  stalls %s" % (
        ie, r[0], 100 * r[2] / toti, 100 * r[1] / tot, {k: round(100 * v / st, 1) for k, v in r[3].most_common(6)}))
    print("      ", ", ".join("%s %d" % kv for kv in r[4].most_common(24)))
if top:
    for a, s, sm, ie, st in sorted(data, key=lambda x: -x[2])[:top]:
        print(hex(a - base), ie, sm, s[:64], st)
