#!/usr/bin/env python3
"""Stall breakdown of a kernel + sample share per block of N SASS instructions (program order): python tools/ncu_regions.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 400
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
for r in rows[2:]:
    print(r[h.index("Kernel Name")][:100])
    d = [(k, v) for k, v in zip(h, r) if ("issue_stalled" in k and k.endswith("per_issue_active.ratio")) or k in (
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")]
    for k, v in sorted(d, key=lambda kv: -float(kv[1].replace(",", "") or 0)):
        print("  ", k.replace("smsp__average_warps_issue_stalled_", ""), v)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; ix = {x: i for i, x in enumerate(hdr)}; data = rows[2:]
stalls = [x for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
for b in range(0, len(data), N):
    blk = data[b:b + N]
    s = sum(int(r[ix["# Samples"]] or 0) for r in blk)
    st = {x[6:]: sum(int(r[ix[x]] or 0) for r in blk) for x in stalls}
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in blk)
    print(blk[0][ix["Address"]][-5:], f"samples {100 * s / tot:5.1f}%  executed {ex:9d}", sorted(st.items(), key=lambda kv: -kv[1])[:4])
