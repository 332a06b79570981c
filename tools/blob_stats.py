"""Print the packer's structure and bank-conflict model for a config (CPU only)."""
import os, sys, struct, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pychain_b200 import synth, engine, ChainGraphBatch
NAMES = ["alpha", "beta'", "gamma", "class", "beta'(cls)"]
def stats(cfg, variant="uniform", B=None):
    c = synth.CONFIGS[cfg]
    g = synth.den_graph_for(cfg, variant=variant)
    gb = ChainGraphBatch(g, 4)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    engine._prepared_cache.clear()
    t = time.time(); prep = engine.prepare_den_graph(gt, g.num_states, c["D"], "cpu", batch_size=B or c["B"]); dt = time.time() - t
    h = prep.blob_host.numpy()
    hdr = struct.unpack_from("<II14i", h, 0)
    opt_g, C, C_pad, R_pad, max_sub, embed = hdr[10], hdr[11], hdr[12], hdr[13], hdr[8], hdr[14]
    off = 64; out = []
    for i in range(5):
        v = struct.unpack_from("<4i5q", h, off); off += 56; out.append(v)
    tail = struct.unpack_from("<6q10d", h, off)
    print(f"{cfg} {variant} G={opt_g}: pack {dt:.2f}s; classes C={C} (R_pad={R_pad}), max_sub={max_sub}, {'embedded row codes' if embed else 'compact streams'}, blob {tail[5]/1e6:.2f} MB")
    tot_wf = 0
    for i in range(5):
        wf, ideal = tail[6 + i], tail[11 + i]
        tot_wf += wf
        print(f"   view{i} {NAMES[i]:11s}: rows {out[i][0]} slices {out[i][1]} split {out[i][2]} recs {out[i][4]} "
              f"wavefronts {wf:.0f} (ideal {ideal:.0f}" + (f", x{wf/ideal:.3f})" if ideal else ")"))
    print(f"   modelled gather wavefronts per frame pair (class view counted once: its sums are spilled): {tot_wf:.0f}")
if __name__ == "__main__":
    for a in (sys.argv[1:] or ["cfg4"]):
        cfg, _, var = a.partition(":")
        stats(cfg, var or "uniform")
