#!/usr/bin/env python
"""bench.py -- LF-MMI chain-loss forward-backward throughput (frames/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config cfg4]

One "step" = one denominator forward-backward (objf + full gradient) over one
batch of synthetic (B,T,D) nnet outputs on a random denominator FST -- the hot
path BASELINE.json names.  Default workload = the configuration the target is
quoted on: cfg4 (S=5000 states, K=40000 arcs, B=512, T=200, D=1000), per GPU.
With N GPUs every rank owns its own B=512 shard (weak scaling; utterances are
independent) and the only collective is one NCCL sum of {objf, frames, not-ok}.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM;
`e2e` = the same metric through the public API (`ChainFunction`) from pinned HOST
buffers (H2D of x and D2H of the objective inside the timed region); `roofline`
relates the dominant kernel to the measured HBM peak via the algorithmic bytes
F*(12D+8S) of SURVEY.md 8(d); `cpu_baseline` = the reference's own CPU path
(oracle/_ref, built from /root/reference by oracle/build_ref.py) timed on this
box's host cores on a bounded sample.

`--impl reference` times that reference CPU path as its own arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LF-MMI frames/sec (fwd+bwd)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--variant", default="uniform", choices=["uniform", "skew", "biphone"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="utterances in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [s.strip() for s in line.split(",")]
            if len(f) >= 7:
                self.samples.append(f)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for f in self.samples:
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ workload
def make_workload(cfg, variant, rank):
    import torch
    from pychain_b200 import synth, ChainGraphBatch
    c = synth.CONFIGS[cfg]
    graph = synth.den_graph_for(cfg, variant=variant)
    x, lengths = synth.make_inputs(c["B"], c["T"], c["D"], seed=2000 + int(cfg[3:]) + 7919 * rank, variable=c["variable"])
    return c, graph, x, lengths


# ------------------------------------------------------------------ reference CPU arm / cpu_baseline
def mode_of(cfg):
    """cfg2 is the leaky-free denominator through the LOG-domain entry (BASELINE.json configs[1]),
    cfg3 the full ChainLoss (numerator + denominator, forward + backward); the rest are the
    probability-domain denominator."""
    return {"cfg2": "logden", "cfg3": "loss"}.get(cfg, "den")


def _ref_worker(args):
    """One host process: the reference extension's CPU path on a shard."""
    cfg, variant, start, count, reps = args
    import torch
    torch.set_num_threads(1)
    from pychain_b200 import synth, ChainGraph, ChainGraphBatch
    from oracle import chain_oracle as co
    c = synth.CONFIGS[cfg]
    mode = mode_of(cfg)
    x, _ = synth.make_inputs(count, c["T"], c["D"], seed=4000 + start)
    lengths = torch.full((count,), c["T"], dtype=torch.long)
    xc = x.clamp(-30, 30)
    use_ref = co.ref_available()
    if mode == "logden":
        lg = ChainGraph(synth.random_den_fst(c["S"], c["K"], c["D"], seed=1002), log_domain=True)
        lgb = ChainGraphBatch(lg, count).materialize()
    else:
        graph = synth.den_graph_for(cfg, variant=variant)
        gb = ChainGraphBatch(graph, count).materialize()
        ex = xc.exp()
    if mode == "loss":
        nb = synth.make_num_batch(count, c["T"], c["D"], seed=52 + start)
    out = []
    objf = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        if mode == "logden":
            o = co.num_ref(lgb, xc, lengths)[0] if use_ref else co.num_c(lgb, xc, lengths)[0]
            objf = float(o)
        else:
            o = co.den_ref(gb, ex, lengths, 1e-5)[0] if use_ref else co.den_c(gb, ex, lengths, 1e-5)[0]
            objf = float(o)
            if mode == "loss":
                o2 = co.num_ref(nb, xc, lengths)[0] if use_ref else co.num_c(nb, xc, lengths)[0]
                objf -= float(o2)
        out.append(time.perf_counter() - t0)
    return out, objf, "reference" if use_ref else "port"


def cpu_reference_run(cfg, variant, sample, reps, max_workers=None):
    """Time `reps` passes of the reference CPU path over `sample` utterances
    sharded over host processes (its arc loops are single-threaded, so
    processes are the only way it can use more than one core).  Returns
    (list of wall seconds per pass, workers, kind, objf)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, sample, max_workers or 64))
    per = [sample // workers + (1 if i < sample % workers else 0) for i in range(workers)]
    starts = [sum(per[:i]) for i in range(workers)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        res = pool.map(_ref_worker, [(cfg, variant, s, n, reps + 1) for s, n in zip(starts, per) if n > 0])
    kind = res[0][2]
    # pass i takes as long as its slowest worker; drop the first (cold) pass
    walls = [max(r[0][i] for r in res) for i in range(1, reps + 1)]
    return walls, len(res), kind, sum(r[1] for r in res)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pychain_b200 import synth
    c = synth.CONFIGS[a.config]
    per_seq_s = 2.0 * c["K"] * c["T"] / (3.0e7 if mode_of(a.config) == "logden" else 3.4e8)  # BASELINE.md cost model, one core
    cores = min(os.cpu_count() or 1, 64)
    sample = a.cpu_sample or max(cores, min(c["B"], int(cores * 2.0 / per_seq_s) or 1))
    sample = min(sample, c["B"])
    walls, workers, kind, _ = cpu_reference_run(a.config, a.variant, sample, max(1, a.steps))
    frames = sample * c["T"]
    mean = sum(walls) / len(walls)
    val = frames / mean
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": len(walls),
        "warmup": 1, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a.config, a.variant, c), "sample_utterances_per_step": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": workers, "kind": kind,
                         "sample": "%d of %d utterances x T=%d per step, %d host processes" % (sample, c["B"], c["T"], workers)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(cfg, variant, c):
    what = {"den": "den fwd+bwd, leaky=1e-5", "logden": "leaky-free den fwd+bwd through the log-domain entry",
            "loss": "full ChainLoss (per-utterance numerators + den, leaky=1e-5) fwd+bwd, variable T"}[mode_of(cfg)]
    return "%s: %s, S=%d states K=%d arcs (%s random FST), B=%d/GPU T=%d D=%d" % (
        cfg, what, c["S"], c["K"], variant, c["B"], c["T"], c["D"])


# ------------------------------------------------------------------ main arm
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist
    from pychain_b200 import synth, engine, ChainGraphBatch, ChainFunction

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    c, graph, x_host, lengths = make_workload(a.config, a.variant, rank)
    B, T, D, S, K = c["B"], c["T"], c["D"], c["S"], c["K"]
    mode = mode_of(a.config)
    from pychain_b200 import ChainGraph, ChainLoss
    if mode == "logden":
        graph = ChainGraph(synth.random_den_fst(S, K, D, seed=1002), log_domain=True)
    gb = ChainGraphBatch(graph, B)
    gt = tuple(getattr(gb, n) for n in (engine._GRAPH_NAMES_LOG if mode == "logden" else engine._GRAPH_NAMES_PROB))
    x_dev = x_host.to(dev)
    x_pin = x_host.pin_memory()
    frames = int(lengths.sum())
    stats = torch.zeros(3, dtype=torch.float32, device=dev)
    if mode == "loss":
        num_batch = synth.make_num_batch(B, T, D, seed=52 + 1000 * rank)
        crit = ChainLoss(graph, 1e-5, avg=False)
        kernel_name = "den_shared_kernel + chain_generic_kernel<log>"
    else:
        kernel_name = "den_shared_log_kernel" if mode == "logden" else "den_shared_kernel"

    class _R(object):
        pass

    def run_engine(xt):
        """One step on a device-resident input; returns an object with .objf and .ok (device tensors)."""
        if mode == "den":
            return engine.forward_backward_prob(gt, xt, lengths, S, 1e-5, x_is_exp=False, clamp=True)
        if mode == "logden":
            return engine.forward_backward_log(gt, xt, lengths, S, grad_as_log=False, clamp=True)
        xg = xt.detach().requires_grad_(True)
        loss = crit(xg, lengths, num_batch)
        loss.backward()
        r = _R()
        r.objf, r.ok = loss.detach(), ChainFunction.last_ok
        return r

    def run_api(xt):
        """The call a user makes (public API) on a staged input; returns the step's scalar result."""
        if mode == "loss":
            return run_engine(xt).objf
        return ChainFunction.apply(xt, lengths, gb, 1e-5) if mode == "den" else ChainFunction.apply(xt, lengths, gb)

    def step_resident():
        r = run_engine(x_dev)
        if world > 1:  # the single collective of the sharded loss: sum of {objf, frames, not ok}
            stats[0] = r.objf
            stats[1] = float(frames)
            stats[2] = (1 - r.ok[0]).float()
            dist.all_reduce(stats)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -------- warm-up (also packs the graph once; not part of a step)
    for _ in range(max(3, a.warmup)):
        r = step_resident()
    barrier()
    launches0 = engine.launch_count()

    # -------- timed: K steps, inputs resident, CUDA events, max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    e0.record()
    for i in range(a.steps):
        kern_ev[i][0].record()
        r = run_engine(x_dev)
        kern_ev[i][1].record()
        if world > 1:
            stats[0] = r.objf
            stats[1] = float(frames)
            stats[2] = (1 - r.ok[0]).float()
            dist.all_reduce(stats)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = engine.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    kern_ms = sum(s.elapsed_time(e) for s, e in kern_ev) / a.steps
    t = torch.tensor([ms_total, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kern_ms = float(t[0]), float(t[1])
    ms_step = ms_total / a.steps
    value = frames * world / (ms_step * 1e-3)
    objf = float(r.objf)
    ok = int(r.ok.item())

    # -------- e2e: public API, pinned host input -> H2D, objf -> D2H, every step.
    # The H2D copy of step i+1 is issued on a copy stream while step i computes (two staging
    # buffers), as a training loop's prefetcher would; every step's copy is inside the timed region.
    stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            stage[i % 2].copy_(x_pin, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def run_e2e(n):
        issue_copy(0)
        last = None
        for i in range(n):
            if i + 1 < n:
                issue_copy(i + 1)  # buffer (i+1)%2 was last read by step i-1, which .item() below has drained
            main_stream.wait_event(ready[i % 2])
            o = run_api(stage[i % 2])
            if world > 1:
                stats[0] = o
                stats[1] = float(frames)
                stats[2] = 0.0
                dist.all_reduce(stats)
                last = stats[0].item()
            else:
                last = o.item()  # D2H of the step's result
        return last

    run_e2e(3)
    barrier()
    e0.record()
    run_e2e(a.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0]) / a.steps
    e2e_val = frames * world / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # -------- roofline of the dominant kernel (den_shared_kernel)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    if os.path.isfile(peaks_path):
        try:
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    alg_bytes = frames * (12.0 * D + 8.0 * S)
    if mode == "loss":
        alg_bytes += frames * (12.0 * D + 8.0 * num_batch.num_states)
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(a.config + ":" + a.variant)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "algorithmic bytes = F*(12D+8S) (x2 for the full loss: numerator + denominator); the kernels are shared-memory-gather / latency bound, see DESIGN.md"}

    # secondary ceiling: the kernels gather 6 fp32 operands per arc and valid frame from
    # shared memory (alpha/x, beta/x, alpha/beta); peak = SMs x 128 B/clk x SM clock
    sm_clock = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    smem_peak = torch.cuda.get_device_properties(dev).multi_processor_count * 128.0 * sm_clock / 1e9
    gather_bytes = 24.0 * K * frames
    roofline["secondary"] = {"bound": "shared-memory gathers", "achieved": gather_bytes / (kern_ms * 1e-3) / 1e9,
                             "peak": smem_peak, "unit": "GB/s",
                             "frac": gather_bytes / (kern_ms * 1e-3) / 1e9 / smem_peak,
                             "bytes_per_launch": gather_bytes}

    cpu_baseline = None
    if not a.no_cpu_baseline and world == 1:
        try:
            per_seq_s = 2.0 * K * T / (3.0e7 if mode == "logden" else 3.4e8)
            cores = min(os.cpu_count() or 1, 64)
            sample = a.cpu_sample or max(cores, min(B, int(cores * 4.0 / per_seq_s) or 1))
            sample = min(sample, B)
            walls, workers, kind, _ = cpu_reference_run(a.config, a.variant, sample, 2)
            w = min(walls)
            cpu_baseline = {"value": sample * T / w, "unit": UNIT, "cores": workers, "kind": kind,
                            "sample": "%d of %d utterances x T=%d, best of 2 passes, %d host processes" % (sample, B, T, workers)}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a.config, a.variant, c), "global_batch": B * world, "frames_per_step": frames * world,
                   "parallelism": "batch-shard x%d, one NCCL sum of {objf,frames,ok}" % world,
                   "l2": "inputs (%.0f MB x + %.0f MB alpha spill per step) exceed the 126 MB L2; no flush needed" % (
                       B * T * D * 4 / 1e6, B * (T + 1) * S * 4 / 1e6)},
        "objf": objf, "ok": ok,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": B * T * D * 4,
                "d2h_bytes_per_step": 4, "api": "pychain_b200.%s; pinned host x -> device every step (copy of step i+1 overlaps step i), objf.item() every step" % ("ChainLoss.forward + backward" if mode == "loss" else "ChainFunction.apply")},
        "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
