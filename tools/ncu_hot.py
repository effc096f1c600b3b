"""Hot instructions of one kernel from an .ncu-rep (source page): python tools/ncu_hot.py rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("samples", tot, "warp-instr", sum(int(r[ix["Instructions Executed"]] or 0) for r in data))
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
    st = sorted(((h[6:], int(r[ix[h]] or 0)) for h in stalls), key=lambda kv: -kv[1])[:2]
    print(r[ix["Address"]][-5:], f"{100*int(r[ix['# Samples']])/tot:5.1f}%", r[ix["Instructions Executed"]], r[ix["Source"]][:64], st)
