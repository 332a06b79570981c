"""Mnemonic counts per kernel of the built library (cuobjdump -sass): what proves the Blackwell-native path.
usage: python tools/sass_evidence.py [lib.so] > profiles/rN_sass_evidence.csv"""
import re, subprocess, sys, os, collections
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pychain_b200", "lib", "libchain_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UBLKCP", "SYNCS", "FFMA2+FMUL2", "LDS.128", "LDS.64", "LDS(32)", "LDG.E.NA.128", "LDG.E.128(L1)", "STS", "BAR", "ATOM/RED", "HMMA/UTC"]
print("# SASS evidence (cuobjdump -sass %s), mnemonic counts per kernel" % os.path.relpath(lib))
print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, FFMA2/FMUL2 = packed fp32x2 (sm_100), LDS.128 = vector shared gathers,")
print("# LDG.E.NA.128 = streaming 128-bit record-pair loads (L1::no_allocate); no tensor-core instructions (UTC*MMA / HMMA) and no atomics by design.")
print("kernel,instructions," + ",".join(cols))
cur, cnt = None, None
def flush():
    if cur:
        print(cur + "," + str(cnt["n"]) + "," + ",".join(str(cnt[c]) for c in cols))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur = re.sub(r"^_ZN9chainb200\d+", "", m.group(1)); cnt = collections.Counter(); continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1); cnt["n"] += 1
    if op.startswith("UBLKCP"): cnt["UBLKCP"] += 1
    elif op.startswith("SYNCS"): cnt["SYNCS"] += 1
    elif op.startswith("FFMA2") or op.startswith("FMUL2"): cnt["FFMA2+FMUL2"] += 1
    elif op.startswith("LDS.128"): cnt["LDS.128"] += 1
    elif op.startswith("LDS.64"): cnt["LDS.64"] += 1
    elif op.startswith("LDS"): cnt["LDS(32)"] += 1
    elif op.startswith("LDG.E.NA.128"): cnt["LDG.E.NA.128"] += 1
    elif op.startswith("LDG.E.128"): cnt["LDG.E.128(L1)"] += 1
    elif op.startswith("STS"): cnt["STS"] += 1
    elif op.startswith("BAR"): cnt["BAR"] += 1
    elif op.startswith("ATOM") or op.startswith("RED"): cnt["ATOM/RED"] += 1
    elif op.startswith("HMMA") or op.startswith("UTC"): cnt["HMMA/UTC"] += 1
flush()
