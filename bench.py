#!/usr/bin/env python
"""bench.py -- ASG forward+backward throughput on B200 (BASELINE.json's metric and config).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|cfg4|cfg5] [--api graph|step|eager-ws|eager]

One "step" = one ASGLoss(L, 'mean') forward + backward over one synthetic batch of the LibriSpeech-like
config (T=1000, N=64 utterances per GPU, L=40, S=200, fp32) -- four scans + the gradient assembly, plus,
for N>1 GPUs, the single fused NCCL all-reduce of the transition gradient and the loss.  Utterances shard
over ranks with no other collective ("scaling": "weak": every GPU keeps 64 utterances, 8 GPUs = BASELINE's
batch-sharded N=512).

Printed JSON line (rank 0): `value` = utterances/s with inputs resident in HBM (CUDA events, max over ranks),
`e2e` = the same metric through the C ABI with HOST buffers (asg_host_loss: H2D of x/A/y/lengths and D2H of
loss, grad_x, grad_A inside the timed region), `roofline` = the gradient-assembly kernel (the HBM-bound kernel
north_star names) timed live with CUDA events on its launching stream, `kernels` = the same for every kernel,
`cpu_baseline` = the compiled unmodified reference (oracle/_ref) on this box's host cores, `clocks` = SM clocks
and throttle reasons sampled during the timed region.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, else the numpy oracle port) on
the same workload; it is the only place besides `cpu_baseline` where anything under oracle/ executes.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.environ.get("ASG_PKG") or os.path.join(ROOT, "torch-asg_b200")      # ASG_PKG: A/B builds (developer)
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: (T, N per GPU, L, S, forward_only)
    "cfg2": (1000, 64, 40, 200, False),     # LibriSpeech-like (BASELINE.json configs[1]) -- the headline
    "cfg3": (4000, 16, 32, 800, True),      # long utterance, forward only (configs[2])
    "cfg4": (500, 128, 1024, 100, False),   # large alphabet: the transition matrix (4 MB) exceeds shared memory (configs[3])
    "cfg5": (1000, 64, 40, 200, False),     # batch-sharded: N=512 over 8 GPUs = cfg2 per GPU (configs[4])
}
METRIC = "ASG fwd+bwd utterances/sec at (T=1000,N=64,L=40)"
L2_BYTES = 126 * 1024 * 1024


def synth(T, N, L, S, seed, dtype):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, L, generator=g).log_softmax(-1).to(dtype)
    A = (0.1 * torch.randn(L, L, generator=g)).to(dtype)
    y = torch.randint(0, L, (N, S), generator=g)
    return x, A, y


def grad_alg_bytes(T, N, L, S):
    """K5 reads alpha, beta, x-equivalent (3 [T,N,L]) + writes gx + reads the two FAC states
    + targets + A in / gA out."""
    return 16 * T * N * L + 8 * T * N * S + 8 * N * S + 8 * L * L


class ClockSampler:
    """The profiling recipe's clocks line: an `nvidia-smi --query-gpu=... -lms` PROCESS started before the warm-up and stopped
    after the timed region; only the samples whose timestamps fall inside the timed region are summarised.  (A pynvml
    thread inside this process was measured to stretch a 0.3 ms graph-replay step to 1-3 ms: every NVML query stalls the
    issuing thread's CUDA calls and takes the GIL from it.  ASG_BENCH_SAMPLER=nvml keeps that variant for comparison.)"""
    FIELDS = "timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period_ms=50):
        import subprocess
        self.proc, self.lines, self.t0, self.t1 = None, [], None, None
        self.mode = os.environ.get("ASG_BENCH_SAMPLER", "smi")
        self.index = index
        self.thread = None
        if self.mode == "smi":
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                              "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self.proc = None
        elif self.mode == "nvml":
            self.thread = _NvmlThread(index)

    def begin(self):
        self.t0 = time.time()
        if self.thread is not None:
            self.thread.start()

    def end(self):
        self.t1 = time.time()
        if self.thread is not None:
            self.thread.stop_flag = True

    def summary(self):
        import datetime
        if self.thread is not None:
            self.thread.join(1.0)
            return self.thread.summary()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": "nvidia-smi not available"}
        time.sleep(0.12)                     # let the sample that covers the end of the region land
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        inside, allrows = [], []
        for ln in out.splitlines():
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 3 + len(self.NAMES):
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                row = (ts, int(f[1]), int(f[2]), [n for n, v in zip(self.NAMES, f[3:]) if v.lower().startswith("active")])
            except Exception:
                continue
            allrows.append(row)
            if self.t0 is not None and self.t0 <= ts <= self.t1 + 0.11:
                inside.append(row)
        rows = inside
        how = "nvidia-smi -lms process, samples inside the timed region"
        if not rows and allrows:             # a timed region shorter than the sampling period: the nearest sample after its start
            after = [r for r in allrows if r[0] >= self.t0]
            rows = after[:1] or allrows[-1:]
            how = "nvidia-smi -lms process, timed region shorter than the sampling period: nearest sample"
        clocks = sorted(r[1] for r in rows)
        reasons = sorted({n for r in rows for n in r[3]})
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": rows[0][2] if rows else None,
                "reasons": reasons, "samples": len(rows), "how": how}


class _NvmlThread(threading.Thread):
    """nvidia-smi's clocks line through NVML from a thread of this process (comparison only, see ClockSampler)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01 if len(self.samples) <= 1 else 0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "how": "pynvml thread"}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation
# --------------------------------------------------------------------------------------------------
def time_reference(T, N, L, S, forward_only, steps, warmup, budget_s=150.0):
    """Returns (seconds per step, utterances per step, kind, cores).  Runs the compiled unmodified reference
    (oracle/_ref, built from /root/reference by oracle/build_ref.py) through the ASGLoss glue on CPU with all
    host threads; falls back to the numpy oracle port if the compiled module did not travel.  Each step is the
    full batch unless `steps` of those would exceed `budget_s`; then a step is the first n utterances of it."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, A, y = synth(T, N, L, S, 0, torch.float32)
    try:
        from oracle import ref_driver
        surf = ref_driver.surface()
        kind = "reference"
        crit = surf["ASGLoss"](L, reduction="mean")
        with torch.no_grad():
            crit.transition.copy_(A)

        def step(n):
            xi = x[:, :n].clone().requires_grad_(True)
            crit.transition.grad = None
            loss = crit(xi, y[:n])
            if not forward_only:
                loss.backward()
            return float(loss.detach())
    except Exception as exc:
        from oracle import asg_oracle as O
        kind, cores = "port", 1
        sys.stderr.write("bench: oracle/_ref unavailable (%s); timing the numpy port\n" % exc)

        def step(n):
            r = O.asg_loss(x[:, :n].numpy(), y[:n].numpy(), A.numpy(), want_grad=not forward_only)
            return float(r["loss"].mean())
    t0 = time.perf_counter()
    step(N)                                             # first touch (also sizes the sample)
    t_full = time.perf_counter() - t0
    n = N
    if t_full * (steps + warmup) > budget_s:
        n = max(1, min(N, int(N * budget_s / (t_full * (steps + warmup)))))
    for _ in range(warmup):
        step(n)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(n)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, n, kind, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, N, L, S, fwd_only = WORKLOADS[args.workload]
    dt, n, kind, cores = time_reference(T, N, L, S, fwd_only, args.steps, args.warmup)
    val = n / dt
    sample = "%d of the batch's %d utterances per step (T=%d,L=%d,S=%d fp32), %d timed steps after %d warm-up, %.2f s/step" % (
        n, N, T, L, S, args.steps, args.warmup, dt)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "utterances/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: T=%d N=%d L=%d S=%d %s, reference CPU path (OpenMP + ATen threads), rank 0 only"
                   % (args.workload, T, N, L, S, "forward-only" if fwd_only else "fwd+bwd")},
        "cpu_baseline": {"value": val, "unit": "utterances/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU baseline)")
    import torch_asg
    import torch_asg_native as nat
    assert nat.IS_B200_BUILD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:
        os.nice(-10)                             # the issuing thread must not lose its core to housekeeping processes
    except OSError:
        pass
    if world > 1:
        # Every NCCL call of this job is issued from one thread, in program order, and waited for on the issuing stream, so a
        # captured all-reduce never overlaps an eager one.  With mixing support on (the default) NCCL makes every launch of a
        # graph that holds a collective wait ON THE HOST for the previous one (measured here: 0.10-0.25 ms per replay vs 0.005).
        os.environ.setdefault("NCCL_GRAPH_MIXING_SUPPORT", "0")
        dist.init_process_group("nccl", device_id=dev)
    T, N, L, S, fwd_only = WORKLOADS[args.workload]
    lib = ctypes.CDLL(os.path.join(PKG, "libasg_b200.so"))
    lib.asg_launch_count.restype = ctypes.c_uint64
    lib.asg_kernel_name.restype = ctypes.c_char_p
    lib.asg_last_error.restype = ctypes.c_char_p

    # inputs larger than L2: rotate over enough distinct batches that a batch is evicted before reuse
    x0, _, y0 = synth(T, N, L, S, 1000 * rank, torch.float32)
    A = synth(T, N, L, S, 0, torch.float32)[1]   # the transition matrix is a shared parameter: the same on every rank
    per_set = x0.numel() * 4
    nsets = max(2, -(-int(1.3 * L2_BYTES) // per_set))
    xs, ys = [], []
    for i in range(nsets):
        xi, _, yi = synth(T, N, L, S, 1000 * rank + i, torch.float32)
        xs.append(xi.to(dev).requires_grad_(not fwd_only))
        ys.append(yi.to(dev))
    il = torch.full((N,), T, dtype=torch.long, device=dev)
    tl = torch.full((N,), S, dtype=torch.long, device=dev)
    crit = torch_asg.ASGLoss(L, reduction="mean", forward_only=fwd_only).to(dev)
    with torch.no_grad():
        crit.transition.copy_(A.to(dev))
    # The public surface with the workspace + CUDA graphs behind it (torch_asg/graphed.py).  Batch-sharded runs pass the
    # process group: the backward graph then ends with the one fused all-reduce [L*L+2].
    gcrit = torch_asg.GraphedASGLoss(crit, group=(dist.group.WORLD if world > 1 else None), global_batch=N * world,
                                     use_graph=(args.api != "eager-ws"), max_graphs=4 * nsets, exchange=args.exchange)
    last = {}

    # Every step hands its input gradient to `last` (what an acoustic model's backward would consume) and clears
    # `x.grad`, so at most two gradient buffers are alive and the caching allocator never goes back to cudaMalloc
    # inside the timed loop (17 retained 10 MB gradients did exactly that in round 1).
    def step_eager(i):                       # the reference-shaped eager route: ASGLoss(...)(...) + backward(), no workspace
        x = xs[i % nsets]
        crit.transition.grad = None
        loss = crit(x, ys[i % nsets], il, tl)
        loss.backward()
        last["loss"], last["gx"] = loss.detach(), x.grad
        x.grad = None

    def step_public(i):                      # THE HEADLINE: loss = criterion(inputs, ...); loss.backward()
        x = xs[i % nsets]
        if fwd_only:
            with torch.no_grad():
                last["loss"] = gcrit(x, ys[i % nsets], il, tl)
            return
        crit.transition.grad = None
        loss = gcrit(x, ys[i % nsets], il, tl)
        loss.backward()
        last["loss"], last["gx"] = loss.detach(), x.grad
        x.grad = None

    def step_onegraph(i):                    # forward + backward + all-reduce as ONE graph replay
        x = xs[i % nsets]
        last["loss"] = gcrit.step(x, ys[i % nsets], il, tl)
        last["gx"] = x.grad                  # the workspace's own buffer: nothing is allocated
        x.grad = None

    if args.api == "eager":
        assert world == 1, "--api eager is the single-GPU reference-shaped route"
    step = {"graph": step_public, "eager-ws": step_public, "step": step_onegraph, "eager": step_eager}[args.api]
    if fwd_only:
        step = step_public

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # construction (not warm-up steps): capture the graphs of every input set, like compiling a model before training
    if args.api != "eager":
        for i in range(nsets):
            gcrit.prepare(xs[i], ys[i], il, tl, kinds=(("fwd_only",) if fwd_only else ("fwd", "bwd", "step")))
    sync_all()

    # ---- NCCL parity self-check (N > 1): rank 0 recomputes the WHOLE batch on its own GPU from the seeds -----------
    parity = None
    if world > 1 and not fwd_only:
        step(0)
        torch.cuda.synchronize()
        g_loss = float(gcrit.global_loss)
        gA = crit.transition.grad.detach().clone()
        gx = last["gx"].detach().clone()
        if rank == 0:
            xw = torch.cat([synth(T, N, L, S, 1000 * r, torch.float32)[0] for r in range(world)], 1).to(dev).requires_grad_(True)
            yw = torch.cat([synth(T, N, L, S, 1000 * r, torch.float32)[2] for r in range(world)], 0).to(dev)
            critw = torch_asg.ASGLoss(L, reduction="mean").to(dev)
            with torch.no_grad():
                critw.transition.copy_(A.to(dev))
            lw = critw(xw, yw)
            lw.backward()
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            parity = {"what": "sharded step (all-reduced) vs the whole batch of %d utterances on rank 0, max-abs-err / max-abs-ref" % (N * world),
                      "loss": abs(g_loss - float(lw)) / abs(float(lw)), "transition_grad": rel(gA, critw.transition.grad),
                      "grad_x_slice": rel(gx, xw.grad[:, :N]), "tolerance": 1e-5}
            parity["ok"] = bool(max(parity["loss"], parity["transition_grad"], parity["grad_x_slice"]) <= parity["tolerance"])
            del xw, yw, critw, lw
        torch.cuda.empty_cache()

    sampler = ClockSampler(local)                # the nvidia-smi process needs ~1 s to start: launch it before the warm-up
    for i in range(args.warmup):
        step(i)
    if args.api in ("graph", "step") and world == 1:   # one eagerly issued step too: the timed loop runs every 8th step that way
        gcrit.use_graph = False                  # (roofline kernel timing) and its one-time costs (side stream of this stream) belong here
        step(args.warmup)
        gcrit.use_graph = True
    sync_all()
    # kernel ids of the library (asg_kernel_name)
    kid = {lib.asg_kernel_name(k).decode(): k for k in range(lib.asg_kernel_count())}
    lib.asg_profile_reset()
    # Roofline kernel, timed LIVE inside the timed region: every PROF_STRIDE-th step runs the same zero-allocation calls
    # eagerly (graph replays cannot carry the library's event pairs) with an event pair around that one kernel on its
    # launching stream.  The other steps are graph replays.
    large = L > 64                               # large alphabets: the FCC scans (a dense product per frame) dominate: fp32-FMA bound
    roof_kernel = "fcc_scan_kernel" if large else ("grad_kernel" if not fwd_only else "fac_scan_kernel")
    lib.asg_profile_select(ctypes.c_uint(1 << kid[roof_kernel]))
    PROF_STRIDE = int(os.environ.get("ASG_BENCH_STRIDE", "8"))      # 0: no eager steps in the timed loop (developer)
    if world > 1:
        # An eagerly issued step also issues its all-reduce eagerly, and NCCL pays ~1-2 ms each time a communicator switches
        # between graph-launched and eager collectives (measured: 2 eager steps stretched a 20-step region from 6.2 to 9.8 ms).
        # Batch-sharded runs therefore keep the timed loop pure replays; the kernels are timed in the separate pass below.
        PROF_STRIDE = 0
    lib.asg_profile_stride(ctypes.c_uint(1))
    lib.asg_profile_enable(1)
    replays0, launches0 = gcrit.graph_replays, gcrit.kernel_launches
    raw0 = lib.asg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    import gc
    gc.collect()
    gc.disable()                                 # the step is issued from Python: keep collector pauses out of the timed loop
    graphed = gcrit.use_graph and args.api != "eager"
    e0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        if graphed and PROF_STRIDE and i % PROF_STRIDE == PROF_STRIDE - 1:
            gcrit.use_graph = False
            step(args.warmup + i)
            gcrit.use_graph = True
        else:
            step(args.warmup + i)
    host_ms = 1e3 * (time.perf_counter() - h0) / max(1, args.steps)      # host time to ISSUE a step (diagnostic)
    e1.record()
    sync_all()
    gc.enable()
    sampler.end()
    ms = e0.elapsed_time(e1)
    launches = int(gcrit.kernel_launches - launches0) if args.api != "eager" else int(lib.asg_launch_count() - raw0)
    replays = int(gcrit.graph_replays - replays0)
    lib.asg_profile_enable(0)
    loss_val = float(last["loss"]) if world == 1 or fwd_only else float(gcrit.global_loss)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.summary()

    def read_kernels():
        ks = {}
        for k in range(lib.asg_kernel_count()):
            tot, cnt = ctypes.c_double(), ctypes.c_uint64()
            lib.asg_profile_read(k, ctypes.byref(tot), ctypes.byref(cnt))
            if cnt.value:
                ks[lib.asg_kernel_name(k).decode()] = {"launches": int(cnt.value), "avg_us": 1e3 * tot.value / cnt.value}
        return ks

    def timed_pass(fn, nsteps, base):
        """A side pass (not part of `value`): ms per step and host issue time per step of another route."""
        for i in range(3):
            fn(base + i)
        sync_all()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.disable()
        p0.record()
        hh = time.perf_counter()
        for i in range(nsteps):
            fn(base + 3 + i)
        hms = 1e3 * (time.perf_counter() - hh) / nsteps
        p1.record()
        sync_all()
        gc.enable()
        tt = torch.tensor([p0.elapsed_time(p1) / nsteps, hms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return {"ms_per_step": float(tt[0]), "host_issue_ms_per_step": float(tt[1])}

    kernels = read_kernels()                     # the roofline kernel, timed inside the timed region
    for v in kernels.values():
        v["note"] = "inside the timed region, every %dth step (eager issue of the same calls)" % PROF_STRIDE
    side = {}
    nside = min(args.steps, 50)
    if not fwd_only and args.api != "eager":
        if args.api != "step":
            side["one_graph_step"] = dict(timed_pass(step_onegraph, nside, 7), api="GraphedASGLoss.step: forward+backward+all-reduce as one graph replay")
        if args.api != "graph":
            side["graphed_autograd"] = dict(timed_pass(step_public, nside, 11), api="GraphedASGLoss(...) + loss.backward()")
        if world == 1:
            side["eager_autograd"] = dict(timed_pass(step_eager, nside, 13), api="ASGLoss(...)(...) + loss.backward(), no workspace, no graph")
    # every kernel, in a separate short pass (not part of `value`): eager issue, event pairs around every launch
    lib.asg_profile_reset()
    lib.asg_profile_select(ctypes.c_uint(0xffffffff))
    lib.asg_profile_stride(ctypes.c_uint(1))
    lib.asg_profile_enable(1)
    gcrit.use_graph = False
    for i in range(min(args.steps, 50)):
        step(args.warmup + args.steps + i)
    gcrit.use_graph = graphed
    sync_all()
    lib.asg_profile_enable(0)
    kernels_all = read_kernels()
    for k, v in kernels_all.items():
        kernels.setdefault(k, dict(v, note="separate pass"))

    api_text = {"graph": "loss = GraphedASGLoss(ASGLoss)(inputs, targets, il, tl); loss.backward() -- the reference's call, workspace + CUDA "
                         "graphs behind it (forward graph, backward graph incl. the all-reduce); x.grad / transition.grad delivered by autograd",
                "eager-ws": "GraphedASGLoss(use_graph=False) + loss.backward(): the zero-allocation C-ABI calls issued eagerly",
                "step": "GraphedASGLoss.step: forward+backward+all-reduce as one graph replay (no autograd graph)",
                "eager": "ASGLoss(...)(...) + loss.backward() (eager autograd, per-call allocations)"}[args.api]
    out = {
        "metric": METRIC if args.workload in ("cfg2", "cfg5") else METRIC + " [workload %s]" % args.workload,
        "value": N * world * args.steps / (ms * 1e-3), "unit": "utterances/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "host_issue_ms_per_step": host_ms,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: T=%d N=%d per GPU (global %d) L=%d S=%d, ASGLoss(mean) %s, full lengths"
                               % (args.workload, T, N, N * world, L, S, "forward-only" if fwd_only else "fwd+bwd"),
                   "l2": "inputs rotate over %d distinct batches (%.0f MB > 126 MB L2); no flush" % (nsets, nsets * per_set / 1e6),
                   "parallelism": ("batch-sharded, one exchange per step: sum over ranks of the fused vector [L*L+2], " + (
                       "one peer-memory kernel per rank over NVLink (asg_peer_allreduce), inside the step's graph" if gcrit.exchange == "peer"
                       else "NCCL all-reduce%s" % (", captured in the graph" if gcrit.nccl_in_graph else ", eager"))) if world > 1 else "single GPU",
                   "api": "ASGLoss forward under no_grad (forward-only graph)" if fwd_only else api_text,
                   "graph_replays": replays, "loss": loss_val, "other_routes": side},
        "gpu_launches": launches, "kernels": kernels, "clocks": clocks,
    }
    if parity is not None:
        out["parity_check"] = parity

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    if large and "fcc_scan_kernel" in kernels:
        # one launch = alpha and beta scans of the whole batch: 2 directions x (T-1) frames x one [L x L] x [L x N] product
        flops = 2.0 * (T - 1) * 2.0 * L * L * N          # 2 directions x (T-1) frames x 2 L^2 N flops
        ach = flops / (kernels["fcc_scan_kernel"]["avg_us"] * 1e-6) / 1e12
        fpeak = 148 * 128 * 2 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e12
        out["roofline"] = {"bound": "fp32_fma", "kernel": "fcc_scan_kernel (resident-tile scan: one dense [LxL]x[LxN] product per frame and direction)",
                           "achieved": ach, "peak": fpeak, "unit": "TFLOP/s", "frac": ach / fpeak, "traffic": None,
                           "algorithmic_flops": flops, "avg_us": kernels["fcc_scan_kernel"]["avg_us"],
                           "peak_source": "nominal fp32 FMA peak: 148 SMs x 128 lanes x 2 x max SM clock (MEASURED_PEAKS.json holds no fp32 figure); "
                                          "tensor cores are not used: the path is specified in fp32"}
    elif "grad_kernel" in kernels:
        alg = grad_alg_bytes(T, N, L, S)
        ach = alg / (kernels["grad_kernel"]["avg_us"] * 1e-6) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "grad_kernel_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        out["roofline"] = {"bound": "hbm", "kernel": "grad_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                           "frac": ach / peak, "traffic": traffic, "algorithmic_bytes": alg, "peak_source": peak_src,
                           "avg_us": kernels["grad_kernel"]["avg_us"]}
    elif "fac_scan_kernel" in kernels:
        x_bytes = 8 * T * N * L
        ach = x_bytes / (kernels["fac_scan_kernel"]["avg_us"] * 1e-6) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": "fac_scan_kernel (forward-only: latency-bound serial scan)", "achieved": ach,
                           "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "algorithmic_bytes": x_bytes,
                           "peak_source": peak_src}

    # ---- end to end through the C ABI with host buffers (rank-local, all ranks concurrently) ----------
    # NCTX host contexts in flight (asg_host_submit / asg_host_wait): every step still copies its inputs from
    # pinned host memory and its results (loss, grad_x, grad_A) back, but step i's H2D overlaps the kernels of the
    # steps before it and the D2H of older ones.  The latency of ONE call (one context, nothing overlapped) is
    # reported next to the pipelined throughput.
    NCTX = int(os.environ.get("ASG_BENCH_NCTX", "8"))
    hx = xs[0].detach().cpu().pin_memory()
    hA = A.clone().pin_memory()
    hy = ys[0].cpu().pin_memory()
    outs = [(torch.empty(N).pin_memory(), torch.empty(T, N, L).pin_memory(), torch.empty(L, L).pin_memory()) for _ in range(NCTX)]
    i64 = ctypes.c_int64
    ctxs = []
    for _ in range(NCTX):
        h = ctypes.c_void_p()
        rc = lib.asg_host_create(0, local, i64(T), i64(N), i64(L), i64(S), ctypes.byref(h))
        assert rc == 0, lib.asg_last_error()
        ctxs.append(h)
    vp = lambda tns: ctypes.c_void_p(tns.data_ptr())

    def e2e_submit(j):
        hloss, hgx, hgA = outs[j]
        rc = lib.asg_host_submit(ctxs[j], vp(hx), vp(hA), vp(hy), None, None, None, i64(T), i64(N), i64(L), i64(S), vp(hloss),
                                 None if fwd_only else vp(hgx), None if fwd_only else vp(hgA))
        assert rc == 0, lib.asg_last_error()

    def e2e_wait(j):
        rc = lib.asg_host_wait(ctxs[j])
        assert rc == 0, lib.asg_last_error()

    def e2e_run(nsteps, nctx):
        for i in range(nsteps):
            if i >= nctx:
                e2e_wait(i % nctx)           # results of step i-nctx are in host memory
            e2e_submit(i % nctx)
        for j in range(min(nctx, nsteps)):
            e2e_wait(j)

    e2e_run(max(3, args.warmup), NCTX)
    sync_all()
    esteps = args.steps
    t0 = time.perf_counter()
    e2e_run(esteps, NCTX)                    # returns when every step's results are in host memory
    dt = time.perf_counter() - t0
    e2e_loss = float(outs[0][0].mean())
    lsteps = min(esteps, 30)
    sync_all()
    t0 = time.perf_counter()
    e2e_run(lsteps, 1)                       # one context: submit, wait, submit, ... = the latency of one call
    lat_ms = 1e3 * (time.perf_counter() - t0) / lsteps
    for h in ctxs:
        lib.asg_host_destroy(h)
    t = torch.tensor([dt, lat_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, lat_ms = float(t[0]), float(t[1])
    h2d = hx.numel() * 4 + hA.numel() * 4 + hy.numel() * 8
    d2h = N * 4 + (0 if fwd_only else (outs[0][1].numel() * 4 + outs[0][2].numel() * 4))
    out["e2e"] = {"value": N * world * esteps / dt, "unit": "utterances/s", "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / esteps, "loss": e2e_loss,
                  "latency_ms_one_call": lat_ms, "latency_utterances_per_s": N * world / (lat_ms * 1e-3),
                  "api": "asg_host_submit/asg_host_wait (C ABI, pinned host buffers, %d contexts in flight; every step copies "
                         "x, A, y in and loss, grad_x, grad_A out); latency_ms_one_call = asg_host_loss-style submit+wait with ONE context" % NCTX}

    # ---- the reference on this box (rank 0, single-GPU run only): its CPU path on the host cores, its CUDA path on this GPU
    if world == 1 and not args.no_cpu_baseline:
        cdt, n, kind, cores = time_reference(T, N, L, S, fwd_only, 5, 1, budget_s=30.0)
        out["cpu_baseline"] = {"value": n / cdt, "unit": "utterances/s", "cores": cores, "kind": kind,
                               "sample": "5 timed steps of %d of the batch's %d utterances (T=%d,L=%d,S=%d fp32) after 1 warm-up, %.2f s/step"
                                         % (n, N, T, L, S, cdt)}
        out["gpu_reference"] = time_reference_gpu(T, N, L, S, fwd_only)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # tear down in order: graphs that hold NCCL kernels first, then the process group; a watchdog ends the process if
        # the teardown itself blocks (a hung exit would otherwise hold the whole box)
        killer = threading.Timer(30.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        gcrit._entries.clear()
        del gcrit
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)


def time_reference_gpu(T, N, L, S, fwd_only, steps=5, warmup=2):
    """The UNMODIFIED reference's CUDA path (its asg.py + its native/*.cpp,*.cu compiled for sm_100a by
    oracle/build_ref.py: build_cuda) on this GPU, in a subprocess that has none of the product on sys.path."""
    import subprocess
    script = os.path.join(ROOT, "oracle", "ref_gpu_bench.py")
    so_dir = os.path.join(ROOT, "oracle", "_ref", "ref_cuda")
    if not (os.path.isdir(so_dir) and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "ref_py"))):
        return {"unavailable": "oracle/_ref/ref_cuda or ref_py did not travel (built by __graft_entry__.build() where /root/reference is mounted)"}
    try:
        r = subprocess.run([sys.executable, script] + [str(v) for v in (T, N, L, S, steps, warmup, int(fwd_only))],
                           capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not line:
            return {"unavailable": "reference CUDA path failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:300]}
        return json.loads(line[-1])
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": "reference CUDA path: %s" % exc}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: how the fused [L*L+2] vector is summed over ranks (auto = the peer-memory kernel when the GPUs have peer access)")
    ap.add_argument("--api", default="step", choices=["graph", "step", "eager-ws", "eager"],
                    help="timed loop: step = GraphedASGLoss.step (default): the graph behind ASGLoss + backward() -- forward, backward and the "
                         "all-reduce as one replay, gradients delivered to inputs.grad / transition.grad; graph = GraphedASGLoss(...)(...) + "
                         "loss.backward() through autograd (forward graph + backward graph); eager-ws = the same zero-allocation calls "
                         "without graphs; eager = plain ASGLoss + backward().  The routes not in the timed loop are timed in short side "
                         "passes and reported under config.other_routes.")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
