#!/usr/bin/env python
"""bench.py - audio samples/s of the TCN/GCN forward on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port of the reference forward)

The contract line (`value`, `e2e`, `roofline`, `cpu_baseline`, `clocks`) is measured on BASELINE.json configs[1]
(cfg2: TCN 10 blocks, 32 ch, k = 15, dilations 2**i, ONE 10 s clip @ 48 kHz mono per GPU); with --gpus N every rank
runs its own clip (weak scaling, no data-path collective; NCCL only broadcasts the weight blob once).  The other
BASELINE configs ride in the same JSON line as sub-records under "configs":

    cfg3  GCN 10 x 32, k = 15, one 10 s clip, the five knob settings (parity of each against the CPU port)
    cfg4  cfg2's network, 64 clips x 10 s per GPU (512 clips at N = 8): device-resident, end to end, parity of >= 8
          full-length clips (first / last of the shards of ranks 0 and N-1 and more inside)
    cfg5  cfg2's network as a stream: 60 min of audio in 65 536-sample chunks and 5 min in 1 024-sample chunks
          (Neutone buffer), per-block history carried on the device; parity of the first 30 s against the CPU port

Prints ONE JSON line on rank 0.  Anything under oracle/ is used only as the checker / CPU baseline.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

SR = 48000
METRIC = "audio samples/sec (48 kHz mono)"
DTYPE = "f32 (3 x fp16 split products on tcgen05, fp32 accumulate; fp32 in / out)"
WORKLOADS = {
    # name: (arch, ctor kwargs, seconds)
    "cfg2": ("TCN", dict(n_channels=32, n_layers=10, dilation_growth=2, kernel_size=15, cond_dim=2), 10.0),
    "cfg3": ("GCN", dict(n_blocks=10, n_channels=32, dilation_growth=2, kernel_size=15, cond_dim=2), 10.0),
    "cfg1": ("TCN", dict(n_channels=16, n_layers=4, dilation_growth=2, kernel_size=3, cond_dim=2), 1.0),
}
CFG3_KNOBS = (0.0, 0.25, 0.5, 0.75, 1.0)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def base_config(workload, world, clips_per_gpu):
    """The `config` dict of the contract line: identical for our arm and the reference arm."""
    arch, kw, seconds = WORKLOADS[workload]
    return dict(workload=workload, arch=arch, **kw, clip_seconds=seconds, clips_per_gpu=clips_per_gpu,
                global_clips=world * clips_per_gpu, sample_rate=SR,
                l2="flushed between timed iterations (256 MiB write)",
                parallelism=f"clips sharded over {world} GPU(s), NCCL weight broadcast only")


def randomise_state(model, seed=0):
    """Default torch init under manual_seed(seed) happened at construction; give BatchNorm
    non-trivial running stats and PReLU slopes so the fold is exercised."""
    g1 = torch.Generator().manual_seed(1)
    g2 = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for blk in model.blocks:
            if hasattr(blk, "film"):
                blk.film.bn.running_mean.copy_(torch.randn(blk.film.bn.running_mean.shape, generator=g1) * 0.5)
                blk.film.bn.running_var.copy_(0.2 + 1.8 * torch.rand(blk.film.bn.running_var.shape, generator=g1))
            if hasattr(blk, "act"):
                blk.act.weight.copy_(0.05 + 0.85 * torch.rand(blk.act.weight.shape, generator=g2))


def build_model(workload):
    import neural_audio_spring_reverb_b200 as N
    arch, kw, seconds = WORKLOADS[workload]
    torch.manual_seed(0)
    if arch == "TCN":
        m = N.TCN(**kw)
    else:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            m = N.GCN(**kw)
    randomise_state(m)
    return m.eval(), arch, kw, int(seconds * SR)


def per_sample_costs(arch, kw):
    """Algorithmic bytes and FLOPs per output sample: per fused block
    4*(Cin+Cout) bytes and 2*Cin*W*k + 2*Cin*Cout FLOPs; last block fused with out_net."""
    C, k = kw["n_channels"], kw["kernel_size"]
    n = kw.get("n_layers", kw.get("n_blocks"))
    W = 2 * C if arch == "GCN" else C
    rows = []
    for i in range(n):
        cin = 1 if i == 0 else C
        cout_bytes = 1 if i == n - 1 else C
        flops = 2 * cin * W * k + 2 * cin * C + (2 * C if i == n - 1 else 0)
        rows.append(dict(bytes=4 * (cin + cout_bytes), flops=flops))
    return rows


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU runs the workload (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.rows, self.proc, self.index, self.enabled = [], None, index, enabled

    def start(self):
        """(Re)start polling.  The sampler only runs around device-timed regions: nvidia-smi polling takes driver locks
        and measurably slows the synchronous host-tensor calls of the end-to-end legs (0.89 vs 1.03 G samples/s at
        -lms 20 on one GPU; with eight ranks on one host every rank feels rank 0's poller)."""
        if not self.enabled or self.proc:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("NASR_SMI_MS", "50")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.proc = None

    def window(self, t_begin, t_end, note):
        """Median SM clock / reasons over the samples taken in [t_begin, t_end] (perf_counter seconds)."""
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no nvidia-smi sample in the window"], samples=0, window=note)
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in list(self.rows):
            if ts < t_begin or ts > t_end:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                power.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    power_w_max=max(power) if power else None, samples=len(sm), window=note)


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(args):
    """The reference's CPU forward (oracle port on the same ATen CPU operators), all host threads."""
    from oracle import nasr_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    model, arch, kw, T = build_model(args.workload)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample: one clip, length chosen so K + W forwards finish in a few minutes
    Ts = min(T, int(args.cpu_seconds * SR))
    x = O.make_input(1, 1, Ts)
    cond = torch.full((1, kw["cond_dim"]), 0.5)
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.warmup, 1))):
        O.forward(sd, model.dilations, x, cond)
    t_one = time.perf_counter() - t0
    # keep the whole arm near a minute whatever K the driver asks for: the forward is an FIR, so a shorter clip of the
    # same workload has the same samples/s (the receptive field is 0.3 s); never below 1 s of audio
    budget_s = 60.0
    if args.steps * t_one > budget_s:
        Ts = max(SR, int(Ts * budget_s / (args.steps * t_one)))
        x = O.make_input(1, 1, Ts)
        O.forward(sd, model.dilations, x, cond)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        O.forward(sd, model.dilations, x, cond)
        times.append(time.perf_counter() - t0)
    mean = statistics.mean(times)
    value = Ts / mean          # mean of K, like the GPU arm (best of K is reported beside it)
    line = dict(impl="reference", metric=METRIC, value=value, unit="samples/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * mean,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=base_config(args.workload, max(world, args.gpus), args.clips_per_gpu),
                note=f"reference forward restated on ATen CPU ops (oracle port); one {Ts / SR:.1f} s clip of the workload per "
                     "step on all host cores, mean of K (the forward is an FIR: samples/s does not depend on the clip length)",
                cpu_baseline=dict(value=value, best_of_k=Ts / min(times), unit="samples/s", cores=cores, kind="port",
                                  sample=f"1 clip x {Ts / SR:.1f} s, fp32, {cores} threads, mean of {args.steps}"),
                e2e=dict(value=value, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm helpers
class Ctx:
    pass


def barrier(ctx):
    if ctx.world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize(ctx.dev)


def max_over_ranks(ctx, v):
    if ctx.world == 1:
        return float(v)
    import torch.distributed as dist
    t = torch.tensor([float(v)], device=ctx.dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def time_device(ctx, model, x, cond, steps, warmup):
    """Device-resident timing: K forwards, each bracketed by CUDA events on the launch stream, L2 flushed in between;
    the device path runs in its asynchronous mode (enqueue only), the fp16-range flag is polled once afterwards.
    Returns (sum of step ms = max over ranks, launches of this rank)."""
    model.set_async(True)
    eng = model._engine()
    for _ in range(warmup):
        model(x, cond)
    barrier(ctx)
    l0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier(ctx)
    t0 = time.perf_counter()
    for s, e in ev:
        ctx.flush.zero_()                      # L2 flush between timed iterations
        s.record()
        model(x, cond)
        e.record()
    barrier(ctx)
    t1 = time.perf_counter()
    launches = eng.launch_count() - l0
    if model.saturated():
        raise RuntimeError("an activation left the fp16 range during the timed region: the timing is void")
    model.set_async(None)
    total_ms = max_over_ranks(ctx, sum(s.elapsed_time(e) for s, e in ev))
    return total_ms, launches, (t0, t1)


def keep_busy(ctx, model, x, cond, t_begin, seconds):
    """Loop the same forward (untimed) until `seconds` have passed since t_begin, so that the clock sampler gets a
    usable window around a timed region that is only milliseconds long.  Returns the end of the window."""
    model.set_async(True)
    while time.perf_counter() - t_begin < seconds:
        for _ in range(4):
            model(x, cond)
        torch.cuda.synchronize(ctx.dev)
    model.set_async(None)
    return time.perf_counter()


def time_e2e(ctx, model, x_host, cond_host, n):
    """End to end through the public API with HOST tensors: every call copies x in, runs, and returns the result in
    pinned host memory (synchronous).  Returns seconds (max over ranks)."""
    y = None
    for _ in range(4):                       # keeps the previous result alive like the timed loop: both pinned blocks get cached
        y = model(x_host, cond_host)
    barrier(ctx)
    t0 = time.perf_counter()
    for _ in range(n):
        y = model(x_host, cond_host)
    torch.cuda.synchronize(ctx.dev)
    return max_over_ranks(ctx, time.perf_counter() - t0), y


def profile_blocks(ctx, model, x, cond, reps=5):
    """Per-block device time in the forward's own pipelined mode (kernel-side %globaltimer stamps, see
    nasr_forward_profiled): block_ms[i] = end_i - end_{i-1}, so the figures add up to the forward's duration."""
    eng = model._engine()
    B, _, T = x.shape
    yb = torch.empty((B, model.out_ch, T), device=ctx.dev)
    stream = torch.cuda.current_stream(ctx.dev).cuda_stream
    eng.set_cond(cond.data_ptr(), B, stream)
    acc = None
    for _ in range(reps):
        ctx.flush.zero_()
        ms = eng.forward_profiled(x.data_ptr(), yb.data_ptr(), B, T, stream)
        acc = ms if acc is None else [a + b for a, b in zip(acc, ms)]
    n = len(acc)
    return [a / reps for a in acc], [eng.block_path(i) for i in range(n)]


def roofline_of(arch, kw, B, T, block_ms, paths, traffic_key=None):
    """Roofline of the dominant kernel = the mid-network block kernel (blocks 1 .. n-1 share it): ALGORITHMIC bytes and
    FLOPs over its average pipelined launch time against the measured peaks.  `frac` is the larger of the
    two algorithmic fractions; the fp16 FLOPs the tensor pipe really executes (3 per algorithmic one) stay beside it."""
    pk = peaks()
    costs = per_sample_costs(arch, kw)
    n = len(block_ms)
    mid = list(range(1, n)) if n > 1 else [0]
    dom_ms = sum(block_ms[i] for i in mid) / len(mid)
    dom_bytes = sum(costs[i]["bytes"] for i in mid) / len(mid) * B * T
    dom_flops = sum(costs[i]["flops"] for i in mid) / len(mid) * B * T
    gbs = dom_bytes / (dom_ms * 1e-3) / 1e9
    tfs = dom_flops / (dom_ms * 1e-3) / 1e12
    tensor_path = all(paths[i] in (1, 2) for i in mid)
    ring_path = sum(1 for i in mid if paths[i] == 2) * 2 > len(mid)
    hbm_frac = gbs / pk["hbm"]
    # ONE tensor denominator: the sustained bf16 / fp16 figure of MEASURED_PEAKS.json (these launches sit inside a long,
    # power-capped sequence of back-to-back kernels)
    tensor_frac = tfs / pk["bf16_sustained"]
    if hbm_frac >= tensor_frac or not tensor_path:
        roof = dict(bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=hbm_frac)
    else:
        roof = dict(bound="tensor", achieved=tfs, peak=pk["bf16_sustained"], unit="TFLOP/s", frac=tensor_frac)
    kernel = ("ring_block_kernel" if ring_path else "tc_block_kernel") if tensor_path else "generic_block_kernel (fp32 FFMA)"
    roof.update(traffic=None, kernel=kernel, peak_source=pk["source"] + "; tensor = sustained bf16 figure",
                algorithmic_bytes_per_launch=dom_bytes, algorithmic_flops_per_launch=dom_flops, launch_ms=dom_ms,
                hbm_gbs=gbs, hbm_frac=hbm_frac, algorithmic_tflops=tfs, tensor_frac_algorithmic=tensor_frac,
                issued_fp16_tflops=3.0 * tfs if tensor_path else None,
                issued_fp16_frac_of_sustained_peak=3.0 * tfs / pk["bf16_sustained"] if tensor_path else None,
                fp32_ffma_frac_of_75tf=tfs / 75.0,
                block_ms=block_ms, block_paths=paths, sum_block_ms=sum(block_ms),
                timing="kernel-side %globaltimer stamps in the forward's own pipelined (PDL) mode; "
                       "block_ms[i] = end_i - end_(i-1), first block from its own start",
                whole_net_hbm_gbs=sum(c["bytes"] for c in costs) * B * T / (sum(block_ms) * 1e-3) / 1e9)
    tpath = ROOT / "profiles" / "traffic.json"
    if tpath.exists():
        try:
            tj = json.loads(tpath.read_text())
            key = traffic_key or kernel.split(" ")[0]
            roof["traffic"] = tj.get(key)
            roof["traffic_note"] = tj.get("_note")
        except Exception:
            pass
    return roof


def cpu_check(model, x_host, cond_host, got, cores, repeats=0):
    """CPU port of the reference forward on the same inputs: parity of `got` (and its timing when repeats > 0)."""
    from oracle import nasr_oracle as O
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    torch.set_num_threads(cores)
    secs = O.time_forward(sd, model.dilations, x_host, cond_host, repeats=repeats, threads=cores) if repeats else None
    ref = O.forward(sd, model.dilations, x_host, cond_host)
    g = got.detach().cpu().double()
    r = ref.double()
    err = float(((g - r).abs().flatten(1).max(1).values / r.abs().flatten(1).max(1).values.clamp_min(1e-30)).max())
    return err, secs


# ------------------------------------------------------------------------------------------------ sub-records
def run_cfg3(ctx, args):
    """GCN + FiLM conditioning sweep: one 10 s clip, five knob settings."""
    model, arch, kw, T = build_model("cfg3")
    model = model.to(ctx.dev).eval()
    g = torch.Generator(device=ctx.dev).manual_seed(300 + ctx.rank)
    x = torch.rand((1, 1, T), device=ctx.dev, generator=g) * 2 - 1
    conds = [torch.full((1, 2), c, device=ctx.dev) for c in CFG3_KNOBS]
    steps = max(5, min(args.steps, 40))
    ctx.sampler.start()
    t_sec0 = time.perf_counter()
    per_knob = []
    total_ms = 0.0
    launches = 0
    for c in conds:
        ms, l, _ = time_device(ctx, model, x, c, steps, 3)
        per_knob.append(T * steps / (ms * 1e-3))
        total_ms += ms
        launches += l
    t_sec1 = keep_busy(ctx, model, x, conds[2], t_sec0, 1.0)
    ctx.sampler.stop()
    value = ctx.world * T * steps * len(conds) / (total_ms * 1e-3)
    x_host = x.cpu().pin_memory()
    e2e_s, _ = time_e2e(ctx, model, x_host, conds[2].cpu().pin_memory(), 20)
    rec = dict(workload="cfg3: GCN 10 x 32, k = 15, dilations 2**i, one 10 s clip, 5 knob settings (c0 = c1)",
               samples_per_s=value, ms_per_clip=total_ms / (steps * len(conds)), steps_per_knob=steps,
               samples_per_s_per_knob=dict(zip(map(str, CFG3_KNOBS), per_knob)), replicas=ctx.world,
               e2e_samples_per_s=ctx.world * T * 20 / e2e_s, gpu_launches=launches)
    if ctx.rank == 0:
        block_ms, paths = profile_blocks(ctx, model, x, conds[2])
        rec["roofline"] = roofline_of(arch, kw, 1, T, block_ms, paths, "ring_block_kernel<gcn>")
        if not args.no_cpu_baseline:
            errs = {}
            secs = None
            for c, cd in zip(CFG3_KNOBS, conds):
                got = model(x, cd)
                err, s = cpu_check(model, x.cpu(), cd.cpu(), got, ctx.cores, repeats=2 if c == 0.5 else 0)
                errs[str(c)] = err
                secs = s or secs
            rec["cpu_baseline"] = dict(value=T / secs, unit="samples/s", cores=ctx.cores, kind="port",
                                       sample="the same 10 s clip, knob 0.5, best of 2",
                                       parity_rel_err_vs_gpu=errs, parity_tolerance=1e-4)
        rec["clocks"] = ctx.sampler.window(t_sec0, t_sec1, "cfg3: the timed forwards + the same forward looped to >= 1 s")
    return rec


def run_cfg4(ctx, args):
    """The batch: 64 clips x 10 s per GPU (BASELINE config 4 = 512 clips over 8 GPUs)."""
    model, arch, kw, T = build_model("cfg2")
    from neural_audio_spring_reverb_b200.distributed import broadcast_weights
    broadcast_weights(model, src=0, device=ctx.dev)
    model = model.to(ctx.dev).eval()
    B = args.cfg4_clips
    g = torch.Generator(device=ctx.dev).manual_seed(400 + ctx.rank)
    x = torch.rand((B, 1, T), device=ctx.dev, generator=g) * 2 - 1
    cond = torch.full((B, 2), 0.5, device=ctx.dev)
    steps = max(3, min(args.steps, 10))
    ctx.sampler.start()
    t_sec0 = time.perf_counter()
    ms, launches, _ = time_device(ctx, model, x, cond, steps, 3)
    t_sec1 = keep_busy(ctx, model, x, cond, t_sec0, 1.0)
    ctx.sampler.stop()
    value = ctx.world * B * T * steps / (ms * 1e-3)
    x_host = x.cpu().pin_memory()
    cond_host = cond.cpu().pin_memory()
    n_e2e = 5
    e2e_s, y_host = time_e2e(ctx, model, x_host, cond_host, n_e2e)
    rec = dict(workload=f"cfg4: cfg2 network, {B} clips x 10 s per GPU, {ctx.world * B} clips over {ctx.world} GPU(s)",
               samples_per_s=value, ms_per_step=ms / steps, steps=steps, clips_per_gpu=B, global_clips=ctx.world * B,
               e2e=dict(value=ctx.world * B * T * n_e2e / e2e_s, unit="samples/s", steps=n_e2e,
                        h2d_bytes_per_step=int(x_host.numel() * 4 + cond_host.numel() * 4),
                        d2h_bytes_per_step=int(y_host.numel() * 4)),
               gpu_launches=launches, scaling="weak")
    # parity: clips {0, B-1} and two inside on ranks 0 and N-1 (N = 1: eight clips on rank 0), full length
    check = []
    if not args.no_cpu_baseline:
        if ctx.world == 1:
            check = sorted({0, 1, B // 4, B // 2 - 1, B // 2, 3 * B // 4, B - 2, B - 1})
        elif ctx.rank in (0, ctx.world - 1):
            check = sorted({0, B // 3, 2 * B // 3, B - 1})
    worst = 0.0
    if check:
        y = model(x, cond)
        for b in check:
            err, _ = cpu_check(model, x[b:b + 1].cpu(), cond[b:b + 1].cpu(), y[b:b + 1], ctx.cores)
            worst = max(worst, err)
        del y
    worst_all = max_over_ranks(ctx, worst)
    if ctx.rank == 0:
        block_ms, paths = profile_blocks(ctx, model, x, cond, reps=2)
        rec["roofline"] = roofline_of(arch, kw, B, T, block_ms, paths, f"ring_block_kernel@B{B}")
        if not args.no_cpu_baseline:
            ranks = [0] if ctx.world == 1 else [0, ctx.world - 1]
            per_rank = 8 if ctx.world == 1 else 4
            rec["parity"] = dict(rel_err_max=worst_all, tolerance=1e-4, clips_checked=per_rank * len(ranks), ranks=ranks,
                                 clips_of_each_rank=check, length="full 10 s", against="CPU port of the reference forward")
        rec["clocks"] = ctx.sampler.window(t_sec0, t_sec1, "cfg4: the timed forwards + the same forward looped to >= 1 s")
    del x, x_host, y_host
    model.release_engine()
    torch.cuda.empty_cache()
    return rec


def run_cfg5(ctx, args):
    """Streaming: chunked causal inference with the per-block history carried across chunks on the device."""
    model, arch, kw, _ = build_model("cfg2")
    model = model.to(ctx.dev).eval()
    cond = torch.full((1, 2), 0.5, device=ctx.dev)
    rec = dict(workload="cfg5: cfg2 network as a stream, history of every block carried on the device", replicas=ctx.world)
    g = torch.Generator(device=ctx.dev).manual_seed(500 + ctx.rank)
    from neural_audio_spring_reverb_b200.streaming import CachedStream
    windows = []
    # 65 536 samples (1.4 s) is the chunk the cfg5 figure is quoted on; 1 024 = a real-time plug-in buffer; 524 288 (10.9 s) shows
    # where chunked processing meets the one-shot forward (the per-chunk cost of every block's warm-up steps is amortised)
    for chunk, seconds in ((65536, args.cfg5_seconds), (1024, min(args.cfg5_seconds, 300.0)), (524288, args.cfg5_seconds)):
        n_chunks = max(4, int(seconds * SR) // chunk)
        total = n_chunks * chunk
        x = torch.rand((1, 1, total), device=ctx.dev, generator=g) * 2 - 1
        stream = CachedStream(model)
        stream.reset(1)
        for j in range(3):                               # warm-up chunks (also sizes the stream planes)
            stream(x[:, :, j * chunk:(j + 1) * chunk], cond)
        stream.reset(1)
        eng = model._engine()
        l0 = eng.launch_count()
        barrier(ctx)
        ctx.sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        keep = []
        keep_n = (30 * SR + chunk - 1) // chunk if chunk == 65536 else 0     # first 30 s kept for the parity check
        for j in range(n_chunks):
            y = stream(x[:, :, j * chunk:(j + 1) * chunk], cond)
            if j < keep_n:
                keep.append(y)
        ev1.record()
        torch.cuda.synchronize(ctx.dev)
        wall = time.perf_counter() - t0
        windows.append((t0, time.perf_counter()))
        ctx.sampler.stop()
        dev_s = max_over_ranks(ctx, ev0.elapsed_time(ev1) * 1e-3)
        wall = max_over_ranks(ctx, wall)
        r = dict(chunk_samples=chunk, audio_seconds=total / SR, chunks=n_chunks,
                 samples_per_s=ctx.world * total / dev_s, us_per_chunk=dev_s / n_chunks * 1e6,
                 rtf=dev_s / (total / SR), wall_samples_per_s=ctx.world * total / wall,
                 launches_per_chunk=(eng.launch_count() - l0) / n_chunks,
                 saturated=bool(model.saturated()))
        if keep and ctx.rank == 0 and not args.no_cpu_baseline:
            got = torch.cat(keep, dim=2)
            n30 = got.shape[2]
            err, secs = cpu_check(model, x[:, :, :n30].cpu(), cond.cpu(), got, ctx.cores, repeats=1)
            r["parity"] = dict(rel_err=err, tolerance=1e-4, seconds=n30 / SR,
                               against="CPU port, one-shot forward over the first 30 s (chunked == one-shot in the reference)")
            rec["cpu_baseline"] = dict(value=n30 / secs, unit="samples/s", cores=ctx.cores, kind="port",
                                       sample=f"one-shot forward over the first {n30 / SR:.1f} s")
        rec[f"chunk_{chunk}"] = r
        del x, keep
    if ctx.rank == 0:
        rec["clocks"] = ctx.sampler.window(windows[0][0], windows[-1][1], "cfg5: the two timed streams")
    model.release_engine()
    torch.cuda.empty_cache()
    return rec


def run_shipped(ctx, args):
    """The reference's own shipped checkpoints (weights in tests/golden): one 10 s clip at the checkpoint's sample rate,
    device-resident, on the kernels the engine picks by default next to the fp32 FFMA kernels (NASR_PATH=fp32), with
    the parity of the default path against the golden vector the reference produced."""
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent / "tests"))
    from util import build_model as build_from_state, golden_inputs, golden_names, load_golden, rel_err
    rec = dict(workload="shipped checkpoints: 1 clip x 10 s each, device-resident, L2 flushed; Msamples/s default path vs "
                        "NASR_PATH=fp32 (FFMA kernels); block_path 0 = FFMA, 1 = tcgen05 tap gather, 2 = tcgen05 ring, "
                        "3 = ring in tap passes", replicas=ctx.world, checkpoints={})
    old = os.environ.get("NASR_PATH")
    for name in golden_names():
        if not name.startswith("ckpt_") or name.endswith("_cond"):
            continue
        meta, y_ref, sd = load_golden(name)
        cfg = meta["cfg"]
        T = 10 * cfg.get("sample_rate", SR)
        g = torch.Generator(device=ctx.dev).manual_seed(7)
        x = torch.rand((1, 1, T), device=ctx.dev, generator=g) * 2 - 1
        cond = torch.tensor([[0.3, 0.7]], device=ctx.dev) if cfg["cond_dim"] else None
        gx, gc = golden_inputs(meta)
        r = dict(arch=cfg["arch"], channels=cfg["n_channels"], kernel_size=cfg["kernel_size"], dilations=meta.get("dilations"))
        for mode in ("auto", "fp32"):
            os.environ["NASR_PATH"] = mode
            m = build_from_state(cfg, sd, ctx.dev)
            m.set_async(True)
            if mode == "auto":
                r["parity_vs_reference_golden"] = rel_err(m(gx.to(ctx.dev), None if gc is None else gc.to(ctx.dev)), y_ref)
                r["block_paths"] = [m._engine().block_path(i) for i in range(len(meta.get("dilations") or [0] * cfg["n_blocks"]))]
            for _ in range(3):
                m(x, cond)
            ms = []
            for _ in range(5):
                ctx.flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                m(x, cond)
                e1.record()
                torch.cuda.synchronize(ctx.dev)
                ms.append(e0.elapsed_time(e1))
            r["msamples_per_s" if mode == "auto" else "msamples_per_s_fp32"] = round(T / sorted(ms)[2] / 1e3, 1)
            m.release_engine()
        rec["checkpoints"][name[5:]] = r
    if old is None:
        os.environ.pop("NASR_PATH", None)
    else:
        os.environ["NASR_PATH"] = old
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # ~0.1 s timed region: long enough for nvidia-smi to see it
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--clips-per-gpu", type=int, default=1)
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5,shipped", help="sub-records to add to the line ('' = none)")
    ap.add_argument("--cfg4-clips", type=int, default=64, help="clips per GPU of the cfg4 sub-record")
    ap.add_argument("--cfg5-seconds", type=float, default=3600.0, help="stream length of the cfg5 sub-record")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="clip length of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    if args.impl == "reference":
        return cpu_reference_arm(args)

    import torch.distributed as dist
    ctx = Ctx()
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path)")
    torch.cuda.set_device(local)
    ctx.dev = torch.device("cuda", local)
    ctx.cores = os.cpu_count() or 1
    if ctx.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=ctx.dev)

    from neural_audio_spring_reverb_b200.build import build_native
    if ctx.rank == 0:
        build_native()
    if ctx.world > 1:
        dist.barrier()
    # CPU pinning per rank is available (NASR_PIN=1) but off by default: measured on an 8-GPU box it does not help the
    # synchronous host-tensor calls (profiles/r2_e2e_scale_8gpu.log: 7.79 G samples/s unpinned vs 7.44 G pinned)
    ctx.affinity = dict(pinned=False, reason="off by default (NASR_PIN=1 enables)")
    if os.environ.get("NASR_PIN", "0") == "1":
        from neural_audio_spring_reverb_b200 import hostaffinity
        ctx.affinity = hostaffinity.pin_to_gpu(local, ctx.world)

    model, arch, kw, T = build_model(args.workload)
    # one-time weight broadcast over NCCL: every rank ends up with rank 0's blob
    from neural_audio_spring_reverb_b200.distributed import broadcast_weights
    broadcast_weights(model, src=0, device=ctx.dev)
    model = model.to(ctx.dev).eval()
    eng = model._engine()

    B = args.clips_per_gpu
    g = torch.Generator(device=ctx.dev).manual_seed(100 + ctx.rank)
    x = torch.rand((B, 1, T), device=ctx.dev, generator=g) * 2 - 1
    cond = torch.full((B, kw["cond_dim"]), 0.5, device=ctx.dev)
    x_host = x.cpu().pin_memory()
    cond_host = cond.cpu().pin_memory()
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)   # > 126 MB L2

    ctx.sampler = ClockSampler(local, enabled=ctx.rank == 0)
    ctx.sampler.start()
    time.sleep(0.15)           # nvidia-smi start-up: the first sample should precede the timed steps

    # ---- device-resident timing (value) ----
    t_main0 = time.perf_counter()
    total_ms, launches, timed = time_device(ctx, model, x, cond, args.steps, args.warmup)
    value = ctx.world * B * T * args.steps / (total_ms * 1e-3)
    # the same forward with the default (checked) device path: every call waits for its result and reads the fp16-range flag
    chk_ms = 0.0
    if ctx.rank == 0:
        n_chk = min(args.steps, 50)
        torch.cuda.synchronize(ctx.dev)
        t0 = time.perf_counter()
        for _ in range(n_chk):
            model(x, cond)
        torch.cuda.synchronize(ctx.dev)
        chk_ms = (time.perf_counter() - t0) * 1e3 / n_chk
    # the timed region of this workload is tens of milliseconds: keep the identical forward running (untimed) until the
    # clock sampler has a usable window around it
    model.set_async(True)
    while time.perf_counter() - t_main0 < 1.5:
        for _ in range(8):
            model(x, cond)
        torch.cuda.synchronize(ctx.dev)
    model.set_async(None)
    t_main1 = time.perf_counter()
    ctx.sampler.stop()         # not during the end-to-end leg (see ClockSampler.start)

    # ---- end to end through the public API with host buffers ----
    # every call is synchronous (H2D, forward, result written to the pinned output, stream sync); at least 50 calls so
    # that a single host hiccup (allocator, scheduler) does not decide a short run
    n_e2e = max(args.steps, 50)
    e2e_s, y_host = time_e2e(ctx, model, x_host, cond_host, n_e2e)
    e2e_value = ctx.world * B * T * n_e2e / e2e_s

    line = None
    if ctx.rank == 0:
        clocks = ctx.sampler.window(t_main0, t_main1, "warm-up + the K timed steps + the same forward looped to >= 1.5 s "
                                    "(nvidia-smi -lms 50)")
        clocks["samples_in_timed_steps"] = ctx.sampler.window(timed[0], timed[1], "")["samples"]
        block_ms, paths = profile_blocks(ctx, model, x, cond)
        roof = roofline_of(arch, kw, B, T, block_ms, paths)
        cpu = None
        if not args.no_cpu_baseline:
            Ts = min(T, int(args.cpu_seconds * SR))
            xs = x_host[:1, :, :Ts].clone()
            cs = cond_host[:1].clone()
            got = model(x[:1, :, :Ts], cond[:1])
            perr, secs = cpu_check(model, xs, cs, got, ctx.cores, repeats=3)   # doubles as a full-size parity check
            cpu = dict(value=Ts / secs, unit="samples/s", cores=ctx.cores, kind="port",
                       sample=f"1 clip x {Ts / SR:.1f} s of the same workload, fp32, {ctx.cores} threads, best of 3",
                       parity_rel_err_vs_gpu=perr, parity_tolerance=1e-4)
        line = dict(metric=METRIC, value=value, unit="samples/s", n_gpus=ctx.world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=total_ms / args.steps,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype=DTYPE, data="synthetic",
                    config=base_config(args.workload, ctx.world, B),
                    device_path="asynchronous (set_async: enqueue only; fp16-range flag polled once after the timed region)",
                    checked_device_path_ms_per_step=chk_ms,
                    e2e=dict(value=e2e_value, unit="samples/s", steps=n_e2e,
                             h2d_bytes_per_step=int(x_host.numel() * 4 + cond_host.numel() * 4),
                             d2h_bytes_per_step=int(y_host.numel() * 4)),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu,
                    host_affinity=ctx.affinity)
    # ---- the other BASELINE configs ----
    del x_host, y_host
    model.release_engine()
    subs = {}
    want = [c for c in args.configs.split(",") if c]
    runners = dict(cfg3=run_cfg3, cfg4=run_cfg4, cfg5=run_cfg5, shipped=run_shipped)
    for name in want:
        if name not in runners:
            raise SystemExit(f"unknown sub-record {name!r}")
        try:
            subs[name] = runners[name](ctx, args)
        except Exception as exc:   # a failing sub-record must not cost the contract line
            subs[name] = dict(error=f"{type(exc).__name__}: {exc}")
            if ctx.world > 1:
                raise
    ctx.sampler.stop()
    if ctx.rank == 0:
        line["configs"] = subs
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
