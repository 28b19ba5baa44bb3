#!/usr/bin/env python
"""bench.py -- ASG forward+backward throughput on B200 (BASELINE.json's metric and config).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|cfg5]

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


class ClockSampler(threading.Thread):
    """nvidia-smi's clocks line through NVML, sampled 10 ms into the timed region and then every 50 ms."""

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
            time.sleep(0.01 if len(self.samples) <= 1 else 0.05)     # every wake-up takes the GIL from the issuing thread

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


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
    from torch_asg.sharded import sharded_asg_step
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
        dist.init_process_group("nccl", device_id=dev)
    T, N, L, S, fwd_only = WORKLOADS[args.workload]
    lib = ctypes.CDLL(os.path.join(PKG, "libasg_b200.so"))
    lib.asg_launch_count.restype = ctypes.c_uint64
    lib.asg_kernel_name.restype = ctypes.c_char_p
    lib.asg_last_error.restype = ctypes.c_char_p

    # inputs larger than L2: rotate over enough distinct batches that a batch is evicted before reuse
    x0, A, y0 = synth(T, N, L, S, 1000 * rank, torch.float32)
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
    last = {}

    from torch_asg.sharded import FusedAllReduce
    fused = FusedAllReduce(crit)

    def step_autograd(i):                    # the plain public API: ASGLoss(...)(...) + backward()
        x = xs[i % nsets]
        x.grad = None
        crit.transition.grad = None
        loss = crit(x, ys[i % nsets], il, tl)
        loss.backward()
        last["loss"] = loss.detach()

    def step(i):
        x = xs[i % nsets]
        if fwd_only:
            with torch.no_grad():
                last["loss"] = crit(x, ys[i % nsets], il, tl)
            return
        if args.api == "autograd" and world == 1:
            return step_autograd(i)
        # One code path at every N: the native forward + backward of this rank's utterances, the mean loss, and (N > 1)
        # the one fused all-reduce of (gA, loss sum, count).  It issues in ~0.1 ms of host time, so the timed loop is bound
        # by the GPU; the autograd route through ASGLoss (same kernels, ~0.2 ms of host time per step, which a busy host
        # stretches past the GPU's 0.32 ms) is timed in a short extra pass and reported as config.public_api_ms_per_step.
        x.grad = None
        last["loss"] = sharded_asg_step(crit, x, ys[i % nsets], il, tl, global_batch=N * world, fused=fused)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    # kernel ids of the library (asg_kernel_name)
    kid = {lib.asg_kernel_name(k).decode(): k for k in range(lib.asg_kernel_count())}
    sampler = ClockSampler(local)
    lib.asg_profile_reset()
    # the timed region carries event pairs only around the kernel whose roofline is reported (two records per step on
    # the main stream); event pairs around the two concurrent scan kernels would add ~270 us of gaps per step
    roof_kernel = "grad_kernel" if not fwd_only else "fac_scan_kernel"
    lib.asg_profile_select(ctypes.c_uint(1 << kid[roof_kernel]))
    PROF_STRIDE = 8                              # an event pair costs the stream ~20 us of launch overlap: time every 8th launch
    lib.asg_profile_stride(ctypes.c_uint(PROF_STRIDE))
    lib.asg_profile_enable(1)
    launches0 = lib.asg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    import gc
    gc.collect()
    gc.disable()                                 # the step is issued from Python: keep collector pauses out of the timed loop
    e0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    host_ms = 1e3 * (time.perf_counter() - h0) / max(1, args.steps)      # host time to ISSUE a step (diagnostic)
    e1.record()
    sync_all()
    gc.enable()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = int(lib.asg_launch_count() - launches0)
    lib.asg_profile_enable(0)
    loss_val = float(last["loss"])
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sampler.join(1.0)

    def read_kernels():
        ks = {}
        for k in range(lib.asg_kernel_count()):
            tot, cnt = ctypes.c_double(), ctypes.c_uint64()
            lib.asg_profile_read(k, ctypes.byref(tot), ctypes.byref(cnt))
            if cnt.value:
                ks[lib.asg_kernel_name(k).decode()] = {"launches": int(cnt.value), "avg_us": 1e3 * tot.value / cnt.value}
        return ks

    public_ms = None
    if world == 1 and not fwd_only and args.api == "native":
        nsteps_pub = min(args.steps, 50)
        for i in range(3):
            step_autograd(i)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        p0.record()
        for i in range(nsteps_pub):
            step_autograd(3 + i)
        p1.record()
        torch.cuda.synchronize()
        public_ms = p0.elapsed_time(p1) / nsteps_pub

    kernels = read_kernels()                     # the roofline kernel, timed inside the timed region
    for v in kernels.values():
        v["note"] = "inside the timed region, every %dth launch" % PROF_STRIDE
    # every kernel, in a separate short pass (not part of `value`): same steps, event pairs around every launch
    lib.asg_profile_reset()
    lib.asg_profile_select(ctypes.c_uint(0xffffffff))
    lib.asg_profile_stride(ctypes.c_uint(1))
    lib.asg_profile_enable(1)
    for i in range(min(args.steps, 50)):
        step(args.warmup + args.steps + i)
    sync_all()
    lib.asg_profile_enable(0)
    kernels_all = read_kernels()
    for k, v in kernels_all.items():
        kernels.setdefault(k, dict(v, note="separate pass"))

    out = {
        "metric": METRIC, "value": N * world * args.steps / (ms * 1e-3), "unit": "utterances/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "host_issue_ms_per_step": host_ms,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: T=%d N=%d per GPU (global %d) L=%d S=%d, ASGLoss(mean) %s, full lengths"
                               % (args.workload, T, N, N * world, L, S, "forward-only" if fwd_only else "fwd+bwd"),
                   "l2": "inputs rotate over %d distinct batches (%.0f MB > 126 MB L2); no flush" % (nsets, nsets * per_set / 1e6),
                   "parallelism": "batch-sharded, one fused all-reduce [L*L+2]" if world > 1 else "single GPU",
                   "api": ("ASGLoss forward (no_grad)" if fwd_only else
                           "ASGLoss(...)(...) + loss.backward() (autograd)" if (args.api == "autograd" and world == 1) else
                           "torch_asg.sharded.sharded_asg_step: native fast_asg_gpu_forward + fast_asg_gpu_backward, mean loss, "
                           "gradients delivered to inputs.grad / transition.grad (same path at every N)"),
                   "loss": loss_val},
        "gpu_launches": launches, "kernels": kernels, "clocks": sampler.summary(),
    }
    if public_ms is not None:
        out["config"]["public_api_ms_per_step"] = public_ms       # ASGLoss + backward() through autograd, extra 50-step pass

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    if "grad_kernel" in kernels:
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
    # steps before it and the D2H of older ones.  Each context owns a stream (and its scans' side stream); the
    # scans of several steps share the GPU (they fill ~2 small CTAs per SM), so throughput ends up bound by the
    # PCIe copies (10 MB each way per step) rather than by one step's scan latency.
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

    def e2e_run(nsteps):
        for i in range(nsteps):
            if i >= NCTX:
                e2e_wait(i % NCTX)           # results of step i-NCTX are in host memory
            e2e_submit(i % NCTX)
        for j in range(min(NCTX, nsteps)):
            e2e_wait(j)

    e2e_run(max(3, args.warmup))
    sync_all()
    esteps = args.steps
    t0 = time.perf_counter()
    e2e_run(esteps)                          # returns when every step's results are in host memory
    dt = time.perf_counter() - t0
    e2e_loss = float(outs[0][0].mean())
    for h in ctxs:
        lib.asg_host_destroy(h)
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    h2d = hx.numel() * 4 + hA.numel() * 4 + hy.numel() * 8
    d2h = N * 4 + (0 if fwd_only else (outs[0][1].numel() * 4 + outs[0][2].numel() * 4))
    out["e2e"] = {"value": N * world * esteps / dt, "unit": "utterances/s", "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / esteps, "loss": e2e_loss,
                  "api": "asg_host_submit/asg_host_wait (C ABI, pinned host buffers, %d contexts in flight; every step copies "
                         "x, A, y in and loss, grad_x, grad_A out)" % NCTX}

    # ---- the reference's CPU path on this box's host cores (rank 0, single-GPU run only) --------------
    if world == 1 and not args.no_cpu_baseline:
        cdt, n, kind, cores = time_reference(T, N, L, S, fwd_only, 5, 1, budget_s=30.0)
        out["cpu_baseline"] = {"value": n / cdt, "unit": "utterances/s", "cores": cores, "kind": kind,
                               "sample": "5 timed steps of %d of the batch's %d utterances (T=%d,L=%d,S=%d fp32) after 1 warm-up, %.2f s/step"
                                         % (n, N, T, L, S, cdt)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--api", default="native", choices=["native", "autograd"],
                    help="timed loop through torch_asg.sharded (native boundary, default) or ASGLoss + backward()")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
