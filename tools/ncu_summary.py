"""Summarise an .ncu-rep: key metrics per kernel + weighted SASS opcode histogram + top stall lines.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if filt not in d.get("Kernel Name", ""):
        continue
    print("=== ", d["Kernel Name"][:90])
    for w in WANT:
        if w in d:
            print("  %-82s %s %s" % (w, d[w], units[hdr.index(w)]))
    st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[k])
          for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k]}
    print("  stalls/issue:", {k: round(v, 2) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]})

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
cur, hdr2 = 'all kernels in report', None
per = collections.defaultdict(lambda: [collections.Counter(), 0, []])
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] in ("Function Name", "Kernel Name"):
        cur, hdr2 = r[1], None
        continue
    if hdr2 is None:
        if "Source" in r and "Instructions Executed" in r:
            hdr2 = r
        continue
    if len(r) != len(hdr2):
        continue
    d = dict(zip(hdr2, r))
    try:
        ie = int(d["Instructions Executed"])
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+(\.MOV)?)", d["Source"])
    op = m.group(2) if m else "?"
    per[cur][0][op] += ie
    per[cur][1] += ie
    try:
        per[cur][2].append((int(d["# Samples"] or 0), d["Source"][:70]))
    except (ValueError, KeyError):
        pass
for k, (ops, tot, samples) in per.items():
    print("--- SASS mix", k[:80], "total warp-instr", tot)
    print("   ", ", ".join("%s %.1f%%" % (o, 100 * v / tot) for o, v in ops.most_common(18)))
    ts = sum(s for s, _ in samples) or 1
    print("    top stall-sample instructions:")
    for s, txt in sorted(samples, reverse=True)[:12]:
        print("      %5.1f%%  %s" % (100 * s / ts, txt))
