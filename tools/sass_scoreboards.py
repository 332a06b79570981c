"""Decode the scheduling control bits of a kernel's SASS (cuobjdump -sass prints both 64-bit words of every
instruction) and report which SCOREBOARD each record-stream load (LDG.E.NA.128) signals, and what its first consumer
waits for.  Volta..Blackwell control layout, bits of the high word: stall 41-44, yield 45, write barrier 46-48
(7 = none), read barrier 49-51, wait mask 52-57.

Why: the SELL passes of den_shared.cu alternate two register banks of stream loads.  A consumer of bank B that waits
on a scoreboard bank A's refill (issued just before) also signals, waits a full L2 latency on every iteration --
ptxas sometimes allocates them that way (6 scoreboards; the gathers take four), and the kernel then runs 3-15 % slower
with identical source semantics.  usage: sass_scoreboards.py lib.so [mangled-kernel-substring]"""
import re, subprocess, sys

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "den_shared_kernelILi4ELb0ELb1E"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ins = None, {}
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); ins[cur] = []; continue
    m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", line)
    if m and cur:
        ins[cur].append([int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]); continue
    m = re.search(r"^\s+/\* (0x[0-9a-f]{16}) \*/", line)
    if m and cur and ins[cur] and ins[cur][-1][3] is None:
        ins[cur][-1][3] = int(m.group(1), 16)

def ctrl(hi):
    return dict(stall=(hi >> 41) & 15, wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7, wait=(hi >> 52) & 63)

bad_total = 0
for fn, L in ins.items():
    if want not in fn:
        continue
    print("==", fn)
    loads = [i for i, x in enumerate(L) if "LDG.E" in x[1] and ".128" in x[1] and not x[1].startswith("@")]
    # groups of consecutive stream loads (a bank fill)
    groups, g = [], []
    for i in loads:
        if g and i - g[-1] > 8:
            groups.append(g); g = []
        g.append(i)
    if g: groups.append(g)
    for g in groups:
        regs, bars = [], []
        for i in g:
            m = re.search(r"\s(R\d+),", L[i][1])
            regs.append(int(m.group(1)[1:])); bars.append(ctrl(L[i][3])["wbar"])
        # first consumer of the bank's first register after the group
        r0 = regs[0]
        cons = None
        for j in range(g[-1] + 1, min(len(L), g[-1] + 4000)):
            if re.search(r"\bR%d\b" % r0, L[j][1].split(",", 1)[-1]) and "LDG" not in L[j][1]:
                cons = j; break
        # loads issued between the group and that consumer
        younger = [k for k in loads if cons and g[-1] < k < cons]
        ybars = sorted(set(ctrl(L[k][3])["wbar"] for k in younger))
        waits = ctrl(L[cons][3])["wait"] if cons else 0
        shared = [b for b in set(bars) if b in ybars]
        flag = "  <-- consumer also waits for YOUNGER stream loads" if shared and any((waits >> b) & 1 for b in shared) else ""
        bad_total += bool(flag)
        print("  loads @%05x regs %s scoreboards %s | first use @%s waits mask %s | younger stream loads in between use %s%s" % (
            L[g[0]][0], regs, bars, ("%05x" % L[cons][0]) if cons else "-", format(waits, "06b"), ybars, flag))
print("bad bank pairs:", bad_total)
