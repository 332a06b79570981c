"""Stall samples per CUDA source line of an ncu report (compiled with -lineinfo, captured with --import-source on).
usage: ncu_lines.py report.ncu-rep [N] [file-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else "den_shared.cu"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        cur = r[1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and cur and want in cur and r and r[0].isdigit() and len(r) >= len(hdr) and "# Samples" in hdr:
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        try:
            n = float(r[i_s] or 0)
        except ValueError:
            continue
        a = agg.setdefault(int(r[0]), dict(n=0.0, src=r[1].strip(), ex=0.0, st={}))
        a["n"] += n
        try:
            a["ex"] += float(r[i_e] or 0)
        except ValueError:
            pass
        for k, v in zip(hdr, r):
            if k.startswith("stall_") and "Not Issued" not in k and v:
                a["st"][k[6:]] = a["st"].get(k[6:], 0.0) + float(v)
data = [(a["n"], ln, a["src"], "%.0f" % a["ex"], sorted(((v, k) for k, v in a["st"].items()), reverse=True)[:3])
        for ln, a in agg.items() if a["n"] > 0]
tot = sum(d[0] for d in data)
print("samples in", want, tot)
for n, ln, src, ex, st in sorted(data, reverse=True)[:N]:
    print(f"{100*n/tot:5.1f}% L{ln:<4d} {src[:84]:84s} exec {ex:>10s}  " + " ".join(f"{k}:{100*v/n:.0f}%" for v, k in st))
