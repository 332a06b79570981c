"""Top stalled SASS instructions of an ncu report (source page).  usage: ncu_hot.py rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    try: data.append((float(r[isamp]), idx, r))
    except ValueError: pass
tot = sum(d[0] for d in data)
print("total samples", tot)
for n, idx, r in sorted(data, key=lambda t: -t[0])[:N]:
    top = sorted([(float(r[i] or 0), hdr[i][6:]) for i in stall], reverse=True)[:2]
    print(f"{n/tot*100:5.1f}% #{idx:5d} {r[isrc].strip()[:70]:70s} exec {r[iexec]:>10s} wf {r[iw]}/{r[iwi]} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
