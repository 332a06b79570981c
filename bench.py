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
    ap.add_argument("--variant", default=None, choices=["uniform", "skew", "biphone"],
                    help="structure of the synthetic denominator FST.  biphone (default; BASELINE.json names cfg4 'biphone "
                         "denominator'): states = phone pairs, the pdf of an arc is a function of its destination, as in "
                         "chain graphs.  uniform: every arc draws its pdf at random (the worst case: no (dst,pdf) class "
                         "has two arcs).  skew: uniform pdfs, Zipf in-degrees.  Default: biphone for cfg4 (whose run also times "
                         "the other two), uniform (SURVEY.md 8d) for the other configs.")
    ap.add_argument("--no-other-variants", action="store_true", help="skip the short timing of the other FST variants")
    ap.add_argument("--cpu-sample", type=int, default=0, help="utterances in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference's own CUDA path")
    ap.add_argument("--no-traffic", action="store_true", help="skip the live ncu measurement of DRAM traffic")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's batch per GPU; strong: the config's GLOBAL batch split over the GPUs")
    ap.add_argument("--probe", action="store_true", help=argparse.SUPPRESS)
    a = ap.parse_args()
    if a.variant is None:
        a.variant = "biphone" if a.config == "cfg4" else "uniform"
    return a


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
                self.samples.append((time.time(), f))

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi's start-up (NVML initialisation over every GPU of the box) takes driver locks that stall kernel
        launches of other processes for up to hundreds of ms: a sampler started right before a short timed region made
        host-bound configs (cfg3: 40 steps in 36 ms) read 1.3-8 ms per step instead of 0.9.  So it is started before
        the warm-up and the timed region begins only once it is sampling."""
        t_end = time.time() + timeout
        while self.proc and not self.samples and time.time() < t_end and self.proc.poll() is None:
            time.sleep(0.02)

    def stop(self, t0=None, t1=None):
        """Clocks over the samples that arrived during [t0, t1] (host times bracketing the timed region; one sampling
        period of slack on both sides -- a region shorter than the 100 ms period may contain no sample of its own)."""
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
        picked = [f for (ts, f) in self.samples if t0 is None or (t0 - 0.12 <= ts <= t1 + 0.16)]
        if not picked:
            picked = [f for (_, f) in self.samples]
        for f in picked:
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
def mode_of(cfg):
    """cfg2 is the leaky-free denominator through the LOG-domain entry (BASELINE.json configs[1]),
    cfg3 the full ChainLoss (numerator + denominator, forward + backward); the rest are the
    probability-domain denominator."""
    return {"cfg2": "logden", "cfg3": "loss"}.get(cfg, "den")


def workload_name(cfg, variant, c):
    what = {"den": "den fwd+bwd, leaky=1e-5", "logden": "leaky-free den fwd+bwd through the log-domain entry",
            "loss": "full ChainLoss (per-utterance numerators + den, leaky=1e-5) fwd+bwd, variable T"}[mode_of(cfg)]
    return "%s: %s, S=%d states K=%d arcs (%s synthetic FST), B=%d/GPU T=%d D=%d" % (
        cfg, what, c["S"], c["K"], variant, c["B"], c["T"], c["D"])


def config_dict(a, spec, world):
    """ONE definition of `config` for both arms (the driver compares them key by key)."""
    B, T, D, S = spec["shard"][1], spec["T"], spec["D"], spec["S"]
    return {"workload": workload_name(a.config, a.variant, dict(spec, B=B)), "global_batch": B * world,
            "frames_per_step": frames_of(spec) * (world if a.scaling == "weak" else 1),
            "scaling": a.scaling,
            "parallelism": "batch-shard x%d, one NCCL sum of {objf,frames,ok}" % world,
            "l2": "inputs (%.0f MB x + %.0f MB alpha spill per step) exceed the 126 MB L2; no flush needed" % (
                B * T * D * 4 / 1e6, B * (T + 1) * S * 4 / 1e6)}


def frames_of(spec):
    from pychain_b200 import synth
    return int(synth.make_lengths(spec["B"], spec["T"], spec["x_seed"], spec["variable"]).sum())


def make_spec(a, rank, world):
    """Weak scaling: every rank owns its own batch of the config's size (seeds differ per rank).  Strong scaling: ONE
    global batch (the config's B_global, cfg5: 4096; rank 0's seeds) of which rank r takes the contiguous shard
    [r*B/W, (r+1)*B/W) -- `shard` = (start, count) inside that batch; the summed objf is then the same for every W."""
    from pychain_b200 import synth
    if a.scaling == "strong":
        spec = synth.workload_spec(a.config, a.variant, 0)
        gb = synth.CONFIGS[a.config].get("B_global", spec["B"])
        spec["B"] = gb
        per = gb // world
        spec["shard"] = (rank * per, per)
    else:
        spec = synth.workload_spec(a.config, a.variant, rank)
        spec["shard"] = (0, spec["B"])
    return spec


def median(v):
    v = sorted(v)
    n = len(v)
    return v[n // 2] if n % 2 else 0.5 * (v[n // 2 - 1] + v[n // 2])


# ------------------------------------------------------------------ reference CPU path (the arm and the cpu_baseline leg)
def cpu_sample_size(spec, seconds_per_pass, explicit=0):
    """Utterances per pass: the whole batch when the box's cores finish it in about `seconds_per_pass`
    (BASELINE.md cost model: 2*K*T / 3.4e8 s per utterance and core; 3.0e7 in the log domain)."""
    per_seq_s = 2.0 * spec["K"] * spec["T"] / (3.0e7 if spec["mode"] == "logden" else 3.4e8)
    cores = min(os.cpu_count() or 1, 64)
    n = explicit or max(cores, min(spec["B"], int(cores * seconds_per_pass / per_seq_s) or 1))
    return min(n, spec["B"])


def cpu_reference(spec, sample, passes, grad_for=()):
    """`passes` timed passes (+1 cold pass, dropped) of the reference's CPU path over utterances
    [0, sample) of the workload, one host process per shard.  ONE estimator for both legs: the median
    pass, each pass as long as its slowest process."""
    from oracle import ref_parallel as rp
    r = rp.run(spec, sample=sample, reps=passes, grad_for=grad_for, drop_first=True)
    r["seconds"] = median(r["walls"])
    r["value"] = r["frames"] / r["seconds"]
    r["desc"] = "%d of %d utterances (%d frames) per pass, median of %d passes after 1 cold pass, %d host processes" % (
        r["sample"], spec["B"], r["frames"], passes, r["workers"])
    return r


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    spec = make_spec(a, 0, world)
    sample = cpu_sample_size(spec, 2.0, a.cpu_sample)
    passes, W = max(1, a.steps), max(3, a.warmup)
    r = cpu_reference(spec, sample, passes + W - 1)   # W warm-up passes: 1 cold (dropped inside) + W-1 dropped here
    walls = r["walls"][W - 1:]
    sec = median(walls)
    val = r["frames"] / sec
    r["desc"] = "%d of %d utterances (%d frames) per step, median of %d steps after %d warm-up passes, %d host processes" % (
        r["sample"], spec["B"], r["frames"], len(walls), W, r["workers"])
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(a, spec, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["workers"], "kind": r["kind"], "sample": r["desc"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "objf": r["objf"] - (r["objf_num"] or 0.0) if spec["mode"] == "loss" else r["objf"],
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ DRAM traffic of the dominant kernel, measured live
def measure_traffic(a, kernel_regex):
    """`dram__bytes_read.sum + dram__bytes_write.sum` of ONE launch of the dominant kernel, from an ncu run
    of this same script (probe mode: warm-up launches only, nothing printed).  None if ncu cannot run."""
    import shutil
    import tempfile
    ncu = shutil.which("ncu")
    if not ncu:
        return None, "ncu not found"
    with tempfile.NamedTemporaryFile(suffix=".csv", delete=False) as f:
        log = f.name
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "-k", "regex:" + kernel_regex, "-s", "1", "-c", "1", "--csv", "--log-file", log,
           sys.executable, os.path.abspath(__file__), "--probe", "--config", a.config, "--variant", a.variant,
           "--scaling", a.scaling]
    try:
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=240,
                       env=dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK=os.environ.get("LOCAL_RANK", "0")))
        import csv
        total, unit_scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rows = [r for r in csv.reader(open(log)) if len(r) > 3]
        head = next(r for r in rows if "Metric Name" in r)
        iname, iunit, ival = head.index("Metric Name"), head.index("Metric Unit"), head.index("Metric Value")
        n = 0
        for r in rows:
            if len(r) > ival and r[iname] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(r[ival].replace(",", "")) * unit_scale.get(r[iunit], 1.0)
                n += 1
        return (total if n == 2 else None), ("ncu --metrics dram__bytes_{read,write}.sum -k regex:%s -c 1, same command, this run" % kernel_regex)
    except Exception as ex:
        return None, "ncu probe failed: %r" % (ex,)
    finally:
        try:
            os.unlink(log)
        except OSError:
            pass


# ------------------------------------------------------------------ main arm
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch
    import torch.distributed as dist
    from pychain_b200 import synth, engine, ChainGraphBatch, ChainFunction, ChainLoss
    from pychain_b200 import distributed as pdist

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

    spec = make_spec(a, rank, world)
    graph, num_batch, x_host, lengths = synth.build_workload(spec, *spec["shard"])
    B, T, D, S, K = spec["shard"][1], spec["T"], spec["D"], spec["S"], spec["K"]
    mode = spec["mode"]
    gb = ChainGraphBatch(graph, B)
    gt = tuple(getattr(gb, n) for n in (engine._GRAPH_NAMES_LOG if mode == "logden" else engine._GRAPH_NAMES_PROB))
    x_dev = x_host.to(dev)
    frames = int(lengths.sum())
    if mode == "loss":
        crit = ChainLoss(graph, 1e-5, avg=False)
        kernel_name, kernel_regex = "den_shared_kernel + chain_generic_kernel<log>", None
    else:
        kernel_name = "den_shared_log_kernel" if mode == "logden" else "den_shared_kernel"
        kernel_regex = "den_shared"

    class _R(object):
        pass

    def run_engine(xt, nb=None, ln=None, g=None):
        """One step on a device-resident input; returns an object with .objf and .ok (device tensors)."""
        ln = lengths if ln is None else ln
        g = gt if g is None else g
        if mode == "den":
            return engine.forward_backward_prob(g, xt, ln, S, 1e-5, x_is_exp=False, clamp=True)
        if mode == "logden":
            return engine.forward_backward_log(g, xt, ln, S, grad_as_log=False, clamp=True)
        xg = xt.detach().requires_grad_(True)
        loss = crit(xg, ln, num_batch if nb is None else nb)
        loss.backward()
        r = _R()
        r.objf, r.ok, r.grad, r.stats = loss.detach(), ChainFunction.last_ok, xg.grad, None
        return r

    def run_api(xt):
        """The call a user makes (public API) on a staged input; returns the step's scalar result."""
        if mode == "loss":
            return run_engine(xt).objf
        return ChainFunction.apply(xt, lengths, gb, 1e-5) if mode == "den" else ChainFunction.apply(xt, lengths, gb)

    def reduce_step(r):
        """The single collective of the sharded loss: sum of {objf, frames, not ok} (finalize_kernel leaves the
        three numbers in one buffer: no extra launch, no host-to-device scalar)."""
        if world > 1:
            if getattr(r, "stats", None) is not None:
                return pdist.reduce_result(r)
            return pdist.reduce_stats(r.objf, frames, r.ok)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler starts BEFORE the warm-up (see ClockSampler.wait_first_sample); only the samples taken during
    # the timed region are reported
    sampler = ClockSampler(local)
    if rank == 0 and not a.probe:
        sampler.start()
    # -------- warm-up (also packs the graph once; not part of a step)
    for _ in range(max(3, a.warmup)):
        r = run_engine(x_dev)
        reduce_step(r)
    barrier()
    if a.probe:   # measure_traffic's child: the profiled launch is one of the warm-up launches
        return
    launches0 = engine.launch_count()

    # -------- timed: K steps, inputs resident, CUDA events, max over ranks
    if rank == 0:
        sampler.wait_first_sample()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_host0 = time.time()
    e0.record()
    for i in range(a.steps):
        kern_ev[i][0].record()
        r = run_engine(x_dev)
        kern_ev[i][1].record()
        red = reduce_step(r)
    e1.record()
    barrier()
    clocks = sampler.stop(t_host0, time.time()) if rank == 0 else None
    launches = engine.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    kern_ms = sum(s.elapsed_time(e) for s, e in kern_ev) / a.steps
    t = torch.tensor([ms_total, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kern_ms = float(t[0]), float(t[1])
    ms_step = ms_total / a.steps
    frames_global = frames
    if world > 1:
        ft = torch.tensor([float(frames)], dtype=torch.float64, device=dev)
        dist.all_reduce(ft)
        frames_global = int(ft[0])
    value = frames_global / (ms_step * 1e-3)
    objf = float(r.objf)
    objf_global = float(red[0]) if red is not None else objf
    ok = int(r.ok.item())

    # -------- e2e: public API, pinned host input -> H2D, objf -> D2H, every step.
    # The H2D copy of step i+1 is issued on a copy stream while step i computes (two staging
    # buffers), as a training loop's prefetcher would; every step's copy is inside the timed region.
    x_pin = x_host.pin_memory()
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
                last = pdist.reduce_stats(o, frames, True)[0].item()
            else:
                last = o.item()  # D2H of the step's result
        return last

    run_e2e(3)
    barrier()
    e0.record()
    run_e2e(a.steps)
    e1.record()
    barrier()
    # the PCIe ceiling of e2e: the H2D copy of one step's input alone
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(3):
        stage[0].copy_(x_pin, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_ms = h0.elapsed_time(h1) / 3
    t = torch.tensor([e0.elapsed_time(e1), h2d_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0]) / a.steps
    h2d_ms = float(t[1])
    e2e_val = frames_global / (e2e_ms * 1e-3)
    del stage

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
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)"
        except Exception:
            pass
    alg_bytes = frames * (12.0 * D + 8.0 * S)
    ro_bytes = frames * (8.0 * D + 4.0 * S)
    if mode == "loss":
        alg_bytes += frames * (12.0 * D + 8.0 * num_batch.num_states)
        ro_bytes += frames * (8.0 * D + 4.0 * num_batch.num_states)
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = None, "not measured"
    if kernel_regex and world == 1 and not a.no_traffic:
        traffic, traffic_src = measure_traffic(a, kernel_regex)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "read_only_achieved": ro_bytes / (kern_ms * 1e-3) / 1e9,
                "read_only_frac": ro_bytes / (kern_ms * 1e-3) / 1e9 / peak,
                "kernel": kernel_name, "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "read_only_bytes_per_launch": ro_bytes, "peak_source": peak_src,
                "note": "algorithmic bytes = F*(12D+8S): x read in the alpha and the beta pass, grad written, alpha row "
                        "written + read (x2 for the full loss); read-only = F*(8D+4S) (north_star's 'HBM-read roofline'); "
                        "the kernels are shared-memory-gather bound, see DESIGN.md"}

    # secondary ceiling: the L1 / shared-memory data pipe, 128 B/clk/SM
    sm_clock = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    smem_peak = torch.cuda.get_device_properties(dev).multi_processor_count * 128.0 * sm_clock / 1e9
    gather_bytes = 24.0 * K * frames
    roofline["secondary"] = {"bound": "shared-memory gathers of the arc-per-record formulation (6 fp32 operands per arc and frame)",
                             "achieved": gather_bytes / (kern_ms * 1e-3) / 1e9, "peak": smem_peak, "unit": "GB/s",
                             "frac": gather_bytes / (kern_ms * 1e-3) / 1e9 / smem_peak, "bytes_per_launch": gather_bytes}

    # -------- parity on the bench tensors + the reference's CPU path timed on this box (same estimator as --impl reference)
    cpu_baseline, parity = None, {}
    if not a.no_cpu_baseline and world == 1:
        try:
            sample = cpu_sample_size(spec, 2.0, a.cpu_sample)
            if mode == "loss":
                sample = B
            grad_for = sorted(set(range(0, sample, max(1, sample // 12))) | {sample - 1})
            ref = cpu_reference(spec, sample, 3, grad_for=grad_for)
            cpu_baseline = {"value": ref["value"], "unit": UNIT, "cores": ref["workers"], "kind": ref["kind"],
                            "sample": ref["desc"]}
            if mode == "loss":
                mine, want = objf, -(ref["objf_num"] - ref["objf"])
                gerr = max(float((r.grad[b].cpu() - (torch.from_numpy(ref["grads"][b][0]) - torch.from_numpy(ref["grads"][b][1]))).abs().max())
                           for b in grad_for)
            else:
                mine, want = float(r.logp[:sample].double().sum()), ref["objf"]
                gerr = max(float((r.grad[b].cpu() - torch.from_numpy(ref["grads"][b])).abs().max()) for b in grad_for)
            parity = {"objf_rel_err": abs(mine - want) / max(1e-30, abs(want)), "grad_max_abs_err": gerr,
                      "parity_vs": "%s CPU path, same tensors, objf over utterances [0,%d), grad over %d of them" % (
                          ref["kind"], sample, len(grad_for)),
                      "objf_ref": want, "objf_mine": mine}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)[:300]}

    # -------- the other structures of the synthetic FST, same sizes, same step, short runs (context for the headline)
    other = None
    if world == 1 and mode == "den" and not a.no_other_variants:
        other = {}
        for var in ("uniform", "skew", "biphone"):
            if var == a.variant:
                continue
            try:
                sp2 = synth.workload_spec(a.config, var)
                sp2["B"] = B
                g2 = synth.build_workload(sp2, 0, 1)[0]
                gb2 = ChainGraphBatch(g2, B)
                gt2 = tuple(getattr(gb2, n) for n in engine._GRAPH_NAMES_PROB)
                for _ in range(3):
                    engine.forward_backward_prob(gt2, x_dev, lengths, S, 1e-5, x_is_exp=False, clamp=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(5):
                    r2 = engine.forward_backward_prob(gt2, x_dev, lengths, S, 1e-5, x_is_exp=False, clamp=True)
                e1.record()
                torch.cuda.synchronize()
                ms2 = e0.elapsed_time(e1) / 5
                other[var] = {"ms_per_step": ms2, "value": frames / (ms2 * 1e-3), "unit": UNIT, "steps": 5,
                              "roofline_frac": alg_bytes / (ms2 * 1e-3) / 1e9 / peak, "ok": int(r2.ok.item())}
            except Exception as ex:
                other[var] = {"error": repr(ex)[:200]}

    # -------- the reference's own CUDA path on this GPU, same tensors
    ref_cuda = None
    if not a.no_ref_cuda and world == 1 and mode in ("den", "logden"):
        try:
            ref_cuda = reference_cuda_leg(spec, graph, x_dev, lengths, r, dev)
        except Exception as ex:
            ref_cuda = {"value": None, "error": repr(ex)[:300]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(a, spec, world),
        "objf": objf_global, "ok": ok,
        "objf_rel_err": parity.get("objf_rel_err"), "grad_max_abs_err": parity.get("grad_max_abs_err"), "parity": parity,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": B * T * D * 4,
                "d2h_bytes_per_step": 4,
                "h2d_ms_alone": h2d_ms, "pcie_ceiling": frames_global / (h2d_ms * 1e-3),
                "api": "pychain_b200.%s; pinned host x -> device every step (copy of step i+1 overlaps step i); the step's "
                       "result read back is the 4-byte objective (objf.item()) -- the gradient stays on the device for the "
                       "network's backward, as in training; pcie_ceiling = frames / (H2D of one step's x alone)" % (
                           "ChainLoss.forward + backward" if mode == "loss" else "ChainFunction.apply")},
        "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "reference_cuda": ref_cuda, "other_variants": other,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_cuda_leg(spec, graph, x_dev, lengths, mine, dev):
    """The UNMODIFIED reference extension's CUDA path (chain-computation.cc:118-135, 256-270 ->
    chain-kernels.cu:247-300; compiled for sm_100 by oracle/build_ref.py) on the bench tensors: the number
    north_star asks to sit next to.  The reference re-uploads the (B-fold replicated) graph on every call
    (chain-computation.cc:77-89); that upload is timed separately as well."""
    import torch
    from oracle import chain_oracle as co
    from pychain_b200 import ChainGraphBatch
    if not co.ref_available():
        return {"value": None, "error": "oracle/_ref not built"}
    B, T = spec["B"], spec["T"]
    mode = spec["mode"]
    gbm = ChainGraphBatch(graph, B).materialize()      # what graph.py:101-120 builds on the host every step
    xc = x_dev.clamp(-30, 30)
    xin = xc.exp() if mode == "den" else xc
    call = (lambda: co.den_ref(gbm, xin, lengths, spec["coeff"])) if mode == "den" else (lambda: co.num_ref(gbm, xin, lengths))
    times = []
    out = None
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = call()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    names = ["forward_transitions", "forward_transition_indices", "forward_transition_probs", "backward_transitions",
             "backward_transition_indices", "backward_transition_probs", "initial_probs", "final_probs"] + (
                 ["leaky_probs"] if mode == "den" else [])
    ups = []
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        keep = [getattr(gbm, n).to(dev) for n in names]
        torch.cuda.synchronize()
        ups.append(time.perf_counter() - t0)
        del keep
    sec, up = median(times[1:]), median(ups)
    frames = int(lengths.sum())
    g_ref = out[1].exp() if mode == "logden" else out[1]
    res = {"value": frames / sec, "unit": UNIT, "ms_per_call": sec * 1e3, "graph_upload_ms": up * 1e3,
           "value_without_graph_upload": frames / max(1e-9, sec - up), "B": B,
           "what": "oracle/_ref (unmodified reference, sm_100) CUDA path, wall clock around the call incl. its host loop, "
                   "its per-call H2D of the replicated graph and its syncs; median of 3 warm calls",
           "objf_rel_err_vs_ours": abs(float(out[0]) - float(mine.objf)) / max(1e-30, abs(float(out[0]))),
           "grad_max_abs_err_vs_ours": float((g_ref - mine.grad).abs().max())}
    return res


if __name__ == "__main__":
    main()
