"""Print the packer's structure and bank-conflict model for a config (CPU only)."""
import os, sys, struct, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pychain_b200 import synth, engine, ChainGraphBatch
def stats(cfg, variant="uniform", B=None):
    c = synth.CONFIGS[cfg]
    g = synth.den_graph_for(cfg, variant=variant)
    gb = ChainGraphBatch(g, 4)
    gt = tuple(getattr(gb, n) for n in engine._GRAPH_NAMES_PROB)
    engine._prepared_cache.clear()
    t = time.time(); prep = engine.prepare_den_graph(gt, g.num_states, c["D"], "cpu", batch_size=B or c["B"]); dt = time.time() - t
    h = prep.blob_host.numpy()
    off = 56; out = []
    opt_g = struct.unpack_from('<i', h, 40)[0]
    for i in range(3):
        v = struct.unpack_from("<4i5q", h, off); off += 56; out.append(v)
    tail = struct.unpack_from("<4q6d", h, off)
    print(f"{cfg} {variant} G={opt_g}: pack {dt:.2f}s; " + "; ".join(
        f"view{i}: slices {out[i][1]} split {out[i][2]} recs {out[i][4]} wf {tail[4+i]:.0f}/{tail[7+i]:.0f}={tail[4+i]/tail[7+i]:.3f}" for i in range(3)))
if __name__ == "__main__":
    for a in (sys.argv[1:] or ["cfg4"]):
        cfg, _, var = a.partition(":")
        stats(cfg, var or "uniform")
