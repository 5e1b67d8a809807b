"""Benchmark of the sonar_b200 hot path (driver contract: ONE JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): sonar_euler_ancestral -- Sonar momentum step fused with
Gaussian ancestral noise -- on SDXL latents 8x4x128x128, 30 sampler steps, defaults (momentum 0.95,
history 0.75, NEW mode, eta 1, s_noise 1); denoiser stub x*0.9 outside the timed regions.
One bench "step" = one 30-step sampling run over one batch. Metric: latent noise elements / second
(elements = 30 x 524,288 per batch).

* value     : device-resident throughput. Every sampler step's kernels (one fused step launch; the
              first step also carries the batched Philox statistics of all 29 noise draws) are
              bracketed by CUDA events on the launching stream; between sampler steps the stub
              denoiser runs: FLUSH_PASSES x 256 MiB of zero-fill, which evicts L2 and occupies the GPU
              for ~0.4 ms (a small fraction of a real UNet call), so the host enqueues the next step
              while the device is still "in the UNet" and the intervals hold device time, not Python
              launch latency. ms_per_step = sum of those 30 event intervals, max over ranks
              (`per_rank_ms` lists every rank: mean run, mean first step).
* e2e       : the same run through the public sampler function with HOST buffers: pinned x0 -> H2D,
              30 steps, result D2H, wall clock between device synchronisations.
* roofline  : dominant kernel (sonar_step_fast_philox2_kernel): algorithmic bytes per launch
              (20 B/element: read x, denoised, history; write x', history'; the noise is regenerated
              from the Philox stream in registers) / CUDA-event duration of that launch with cold L2,
              against MEASURED_PEAKS.json hbm_gbs; the same kernel on the C5 per-GPU shard shape is
              reported as roofline_large_tensor.
* cpu_baseline : the CPU oracle port of the reference algorithm (oracle/sonar_oracle.py) on the host
              cores, same workload, bounded sample.
* N > 1     : weak scaling by batch: every rank holds 8 latents of a global batch of 8N; the global
              scale_noise statistics of all 29 draws are summed over ranks ONCE per run (29x2 doubles through
              the NVLink peer mailboxes, inside the first sampler step). Before each run the ranks meet at a
              host barrier and, at the run's first denoiser call, at a device-side rendezvous
              (parallel.device_barrier) behind which every host keeps enqueueing.
"""

from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

SHAPE = (8, 4, 128, 128)
N_SAMPLER_STEPS = 30
FLUSH_PASSES = 8  # 256 MiB zero-fills per stub-denoiser call: L2 eviction + ~0.4 ms of device time (see StepTimer)
ELEMS_PER_RUN = N_SAMPLER_STEPS * SHAPE[0] * SHAPE[1] * SHAPE[2] * SHAPE[3]
WORKLOAD = "C2 sonar_euler_ancestral, SDXL latents 8x4x128x128, 30 steps, fused Gaussian noise"


def make_sigmas() -> torch.Tensor:
    return torch.cat((torch.linspace(14.6, 0.03, N_SAMPLER_STEPS), torch.zeros(1)))


def measured_peak() -> tuple[float, str]:
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device_index: int, enabled: bool = True):
        self.device_index = device_index
        self.enabled = enabled  # rank 0 samples its GPU; N pollers would only add host contention
        self.proc = None
        self.lines: list[str] = []

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            # nvidia-smi spends ~0.1-1 s initialising NVML (and holds driver locks while it does): wait for its
            # first sample so that start-up does not land inside the first timed runs; the 100 ms polling that
            # follows is what samples the timed region
            deadline = time.perf_counter() + 5.0
            while not self.lines and time.perf_counter() < deadline and self.proc.poll() is None:
                time.sleep(0.01)
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm, sm_max, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(sm_max) if sm_max else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(n_runs: int, sampler_steps: int = N_SAMPLER_STEPS) -> dict:
    """Times oracle.SonarOracle.euler_ancestral + CPU Gaussian noise + scale_noise (what the
    reference's SonarEulerAncestral.step does on CPU, py/sonar.py:541-573), all host threads."""
    from oracle import sonar_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sigmas = make_sigmas()
    torch.manual_seed(0)
    x0 = torch.randn(SHAPE) * sigmas[0]
    times = []
    for _ in range(n_runs):
        o = orc.SonarOracle()
        x = x0.clone()
        den = x * 0.9
        t_run = 0.0
        for i in range(sampler_steps):
            t0 = time.perf_counter()
            noise = orc.scale_noise(torch.randn(SHAPE), 1.0, normalized=True) if sigmas[i + 1] > 0 else None
            x = o.euler_ancestral(i, x, den, sigmas[i], sigmas[i + 1], noise)
            t_run += time.perf_counter() - t0
            den = x * 0.9  # denoiser stub, untimed
        times.append(t_run)
    best = min(times)
    elems = sampler_steps * x0.numel()
    return {
        "value": elems / best,
        "unit": "elements/s",
        "cores": cores,
        "kind": "port",
        "sample": f"{n_runs} x {sampler_steps} sampler steps of C2 on CPU (oracle port, best run, stub denoiser untimed)",
        "ms_per_step": best * 1e3 * (N_SAMPLER_STEPS / sampler_steps),
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    for _ in range(args.warmup):
        cpu_reference_run(1, sampler_steps=10)
    res = cpu_reference_run(max(1, args.steps))
    line = {
        "impl": "reference",
        "metric": "latent noise elements/sec",
        "value": res["value"],
        "unit": "elements/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "where": "host CPU", "threads": res["cores"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class StepTimer:
    """Denoiser stub that doubles as the boundary of the timed regions: everything the sampler
    enqueues between two model calls is the hot path of one sampler step."""

    def __init__(self, device, flush_bytes: int = 256 << 20, timed: bool = True, flush_passes: int = FLUSH_PASSES,
                 on_first_call=None):
        self.flush = torch.empty(flush_bytes, dtype=torch.uint8, device=device) if flush_bytes else None
        self.flush_passes = flush_passes
        self.timed = timed
        self.on_first_call = on_first_call  # multi-GPU: device-side rendezvous of the ranks at the first denoiser call
        self.pairs: list[tuple[torch.cuda.Event, torch.cuda.Event]] = []
        self._open: torch.cuda.Event | None = None

    def close(self):
        if self._open is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self.pairs.append((self._open, end))
            self._open = None

    def __call__(self, x, sigma, **_kw):
        if self.timed:
            self.close()
        if self.on_first_call is not None:
            # the GPUs meet HERE, with the host already enqueueing the denoiser and the first sampler step behind
            # the rendezvous: from its release on every rank runs from a full queue, so the exchange inside the
            # first step measures NVLink and the slowest GPU, not the slowest Python thread
            self.on_first_call()
            self.on_first_call = None
        if self.flush is not None:
            # evicts L2 and keeps the GPU busy for ~0.4 ms, like (a small fraction of) the UNet forward that sits
            # here in real use: the host enqueues the next step while the device is still in the denoiser, so
            # the timed intervals are device time of the hot path, not Python launch latency
            for _ in range(self.flush_passes):
                self.flush.zero_()
        den = x * 0.9
        if self.timed:
            self._open = torch.cuda.Event(enable_timing=True)
            self._open.record()
        return den

    def total_ms(self) -> float:
        return sum(a.elapsed_time(b) for a, b in self.pairs)


def step_kernel_roofline(sb, dev, shape, peak: float, peak_src: str, reps: int) -> dict:
    """Times the dominant kernel alone (sonar_step_fast_philox_kernel: momentum mix + both history
    updates + Euler step + ancestral noise regenerated from the Philox stream and normalised from the
    look-ahead statistics). Algorithmic bytes: 20 B/element = read x, denoised, history; write x',
    history' (the noise never touches HBM).

    The launches of several independent operand sets are captured into ONE CUDA graph, so the interval
    between the two CUDA events is device time of back-to-back kernels, not host launch cadence (a
    ctypes launch costs ~5 us on the host, as long as the kernel itself at this size). L2 is flushed
    (256 MiB write) before every replay and each operand set is touched once per replay: cold reads."""
    import statistics as st

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n = 1
    for d in shape:
        n *= d
    n_sets = max(2, min(8, (96 << 20) // (20 * n)))
    sets = []
    for _ in range(n_sets):
        x, den, hist = (torch.randn(shape, device=dev) for _ in range(3))
        draw = sb.ops.reserve_draw(n, dev)
        sums = sb.ops.philox_normal_moments_batch(draw, [draw.offset], begin=0, count=n, device=dev)
        dec = sb.ops.norm_decisions(sums, n)
        stepper = sb.samplers.SonarBase(sb.samplers.SonarConfig())
        kw = {"draw": draw, "factor": 1.0, "normalized": True, "begin": 0, "sums": (sums, dec), "sums_ptr": sums.data_ptr(),
              "decision_ptr": dec.data_ptr(), "count": n}
        sets.append((stepper, x, den, hist, kw))
    outs = []

    def launch_all():
        outs.clear()
        for stepper, x, den, hist, kw in sets:
            stepper.history_d = hist
            outs.append(stepper.fused_step(3, x, den, 5.0, kind=sb.ops.STEP_EULER, c0=-1.5, noise_scale=0.7, noise_philox=kw))

    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        for _ in range(3):
            launch_all()
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        launch_all()
    for _ in range(3):
        flush.zero_()
        graph.replay()
    torch.cuda.synchronize()
    us = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us.append(e0.elapsed_time(e1) * 1e3 / n_sets)
    launch_us = st.median(us)
    algo = 20 * n
    achieved = algo / (launch_us * 1e-6) / 1e9
    # draws of at most two ATen rows (numel <= 2 * 256 * grid) take the single-wave variant
    two_rows = n <= 2 * 256 * sb.ops.philox_policy(n)[0]
    return {
        "bound": "hbm",
        "kernel": "sonar_step_fast_philox2_kernel" if two_rows else "sonar_step_fast_philox_kernel",
        "shape": list(shape),
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo,
        "launch_us": launch_us,
        "launches_timed": len(us) * n_sets,
        "timing": f"CUDA events around one graph replay of {n_sets} back-to-back launches on distinct operand sets, after a 256 MiB L2 flush",
    }


def sampler_run(sb, model, x0, sigmas):
    return sb.samplers.SonarEulerAncestral.sampler(model, x0, sigmas, extra_args={"seed": 0}, disable=True)


def run_b200_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; sonar_b200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import sonar_b200 as sb

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sigmas_host = make_sigmas()
    sigmas = sigmas_host.to(dev)
    torch.manual_seed(1234 + rank)
    x0 = torch.randn(SHAPE, device=dev) * sigmas_host[0]
    x0_host = x0.cpu().pin_memory()
    global_batch = SHAPE[0] * world

    def one_run(model, src):
        torch.manual_seed(99)  # replicated generator: every rank reserves the same global draws
        if world > 1:
            with sb.parallel.sharded(global_batch, rank=rank, world_size=world):
                return sampler_run(sb, model, src, sigmas)
        return sampler_run(sb, model, src, sigmas)

    # ---------------- device-resident throughput ----------------
    warm_model = StepTimer(dev, timed=False)
    def device_rendezvous():
        with sb.parallel.sharded(global_batch, rank=rank, world_size=world):
            sb.parallel.device_barrier()

    rendezvous = device_rendezvous if world > 1 else None

    def align_ranks():
        """Ranks start a run together: a host barrier here, and a device-side rendezvous (peer mailboxes) at the
        run's first denoiser call (StepTimer.on_first_call)."""
        barrier()

    for _ in range(max(3, args.warmup)):
        align_ranks()
        one_run(StepTimer(dev, timed=False, on_first_call=rendezvous), x0)
    barrier()
    launches0 = sb.ops.LAUNCH_COUNT
    timers = []
    gc.collect()
    gc.disable()  # a collection pause between two launches would show up as device idle time
    with ClockSampler(local_rank, enabled=rank == 0) as clocks:
        align_ranks()
        one_run(StepTimer(dev, timed=False, on_first_call=rendezvous), x0)  # one more untimed run with the clock poller up
        barrier()
        t_wall = time.perf_counter()
        for _ in range(args.steps):
            if world > 1:
                align_ranks()
            timer = StepTimer(dev, on_first_call=rendezvous)
            one_run(timer, x0)
            timer.close()
            timers.append(timer)
        barrier()
        wall = time.perf_counter() - t_wall
        # keep the sampler busy long enough for a few clock samples; a FIXED run count, because every
        # sharded run performs one exchange and all ranks must perform the same number of them
        for _ in range(300):
            one_run(warm_model, x0)
        torch.cuda.synchronize()
    gc.enable()
    launches = sb.ops.LAUNCH_COUNT - launches0
    run_ms = [t.total_ms() for t in timers]
    ms_per_step = statistics.mean(run_ms)
    # per rank: (mean run, mean first-step interval -- the one that holds the look-ahead pass and the exchange)
    first_ms = statistics.mean(t.pairs[0][0].elapsed_time(t.pairs[0][1]) for t in timers)
    per_rank = torch.tensor([ms_per_step, first_ms], device=dev, dtype=torch.float64)
    if world > 1:
        gathered = [torch.zeros_like(per_rank) for _ in range(world)]
        dist.all_gather(gathered, per_rank)
        per_rank_ms = [[round(float(v), 4) for v in g.tolist()] for g in gathered]
    else:
        per_rank_ms = [[round(ms_per_step, 4), round(first_ms, 4)]]
    ms_per_step = max(r[0] for r in per_rank_ms) if world > 1 else ms_per_step
    value = ELEMS_PER_RUN * world / (ms_per_step * 1e-3)

    # ---------------- roofline of the dominant kernel (per-launch CUDA events) ----------------
    peak, peak_src = measured_peak()
    roofline = step_kernel_roofline(sb, dev, SHAPE, peak, peak_src, reps=40)
    roofline_large = step_kernel_roofline(sb, dev, (1, 16, 33, 90, 160), peak, peak_src, reps=20)
    traffic_path = REPO / "profiles" / "traffic.json"
    if traffic_path.exists():
        traffic = json.loads(traffic_path.read_text())
        roofline["traffic"] = traffic.get(roofline["kernel"] + "@8x4x128x128")
        roofline_large["traffic"] = traffic.get(roofline_large["kernel"] + "@1x16x33x90x160")

    # ---------------- end to end through the public API with host buffers ----------------
    e2e_times = []
    e2e_model = StepTimer(dev, flush_bytes=0, timed=False)
    gc.collect()
    gc.disable()
    # result lands in a pinned host buffer (what a serving loop does; a pageable destination adds a
    # staging copy and first-touch page faults to every run)
    out_host = torch.empty((global_batch if rank == 0 else SHAPE[0], *SHAPE[1:]), dtype=torch.float32).pin_memory()
    for i in range(3 + args.steps):
        barrier()
        t0 = time.perf_counter()
        x_dev = x0_host.to(dev, non_blocking=True)
        out = one_run(e2e_model, x_dev)
        if world > 1:
            with sb.parallel.sharded(global_batch, rank=rank, world_size=world):
                out = sb.parallel.gather(out, dst=0)
        if out is not None:
            out_host[: out.shape[0]].copy_(out, non_blocking=True)
        torch.cuda.synchronize()
        if i >= 3:
            e2e_times.append(time.perf_counter() - t0)
    gc.enable()
    e2e_s = statistics.mean(e2e_times)
    t_dev = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    e2e_s = float(t_dev.item())
    bytes_x = x0.numel() * 4
    e2e = {
        "value": ELEMS_PER_RUN * world / e2e_s,
        "unit": "elements/s",
        "h2d_bytes_per_step": bytes_x + sigmas.numel() * 0,
        "d2h_bytes_per_step": bytes_x * (world if rank == 0 else 1),
        "ms_per_step": e2e_s * 1e3,
        "note": "public sampler function, pinned host x0 -> H2D, 30 steps (stub denoiser inside), result D2H"
        + (" after NCCL gather to rank 0" if world > 1 else ""),
    }

    extras = None
    cpu_baseline = None
    if rank == 0 and world == 1:
        if not args.no_extras:
            extras = run_extras(sb, dev, peak)
        cpu = cpu_reference_run(3)
        cpu_baseline = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "latent noise elements/sec",
            "value": value,
            "unit": "elements/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "global_batch": global_batch,
                "per_gpu_shape": list(SHAPE),
                "parallelism": f"batch-sharded x{world}" if world > 1 else "single GPU",
                "l2": f"{FLUSH_PASSES} x 256 MiB flush between sampler steps (where the UNet runs, ~0.4 ms of device time); stub denoiser untimed",
                "timing": "sum of 30 CUDA-event intervals per run (fused step launch; step 0 includes the batched Philox statistics of all draws), max over ranks",
            },
            "roofline": roofline,
            "roofline_large_tensor": roofline_large,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "wall_s_timed_region": wall,
            "runs_ms": run_ms,
            "per_rank_ms": per_rank_ms,  # [mean run, mean first sampler step] of every rank
        }
        if extras is not None:
            line["other_configs"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# the other BASELINE.json configs (single GPU): elements/s and roofline fraction of the top kernel
# ---------------------------------------------------------------------------------------------
def _time_traced(sb, fn, reps: int, flush: torch.Tensor) -> tuple[float, dict]:
    """Mean total kernel time (us) and per-kernel means over `reps` traced invocations."""
    for _ in range(3):
        fn()
    t_end = time.perf_counter() + 0.3  # the GPU drops to idle clocks while Python sets a config up
    while time.perf_counter() < t_end:
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    totals, per = [], {}
    for _ in range(reps):
        flush.zero_()
        sb.ops.TRACE = []
        fn()
        torch.cuda.synchronize()
        trace, sb.ops.TRACE = sb.ops.TRACE, None
        tot = 0.0
        for name, a, b in trace:
            us = a.elapsed_time(b) * 1e3
            per.setdefault(name, []).append(us)
            tot += us
        totals.append(tot)
    return statistics.mean(totals), {k: statistics.mean(v) * (len(v) / reps) for k, v in per.items()}


def run_extras(sb, dev, peak: float) -> list[dict]:
    import math

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []

    def record(name, elems, algo_bytes_per_elem, total_us, per, note):
        top = max(per, key=per.get)
        out.append(
            {
                "config": name,
                "elements": elems,
                "kernel_us": total_us,
                "elements_per_s": elems / (total_us * 1e-6),
                "algorithmic_bytes_per_element": algo_bytes_per_elem,
                "hbm_gbs": elems * algo_bytes_per_elem / (total_us * 1e-6) / 1e9,
                "frac_of_measured_peak": elems * algo_bytes_per_elem / (total_us * 1e-6) / 1e9 / peak,
                "top_kernel": top,
                "per_kernel_us": per,
                "note": note,
            },
        )

    # C1: SonarPowerNoise pink on 1x4x64x64 (RNG draw of the half spectrum + irfft2 + normalisation)
    ng, sn = sb.noise_graph, sb.spectral_noise
    power_kw = dict(time_brownian=False, alpha=1.0, max_freq=0.7071, min_freq=0.0, stretch=1.0, rotate=0.0, pnorm=2.0,
                    mix=1.0, common_mode=0.0, channel_correlation="1, 1, 1, 1, 1, 1")  # fmt: skip

    def power_chain():
        c = ng.CustomNoiseChain()
        c.add(sn.PowerNoiseItem(1.0, **power_kw))
        return c

    x = torch.zeros(1, 4, 64, 64, device=dev)
    ns = power_chain().make_noise_sampler(x, None, None, seed=0)
    us, per = _time_traced(sb, lambda: ns(None, None), 10, flush)
    record("C1 SonarPowerNoise pink 1x4x64x64", x.numel(), 8.125 + 8.125 + 12, us, per,
           "bytes: spectrum write+read 2x8.125, irfft2 write 4, moments read 4, scale r/w... (see DESIGN.md)")

    # C3: Scheduled(Blended(lerp .5, pyramid, perlin), fallback gaussian) on 16x16x128x128
    def chain_of(t):
        c = ng.CustomNoiseChain()
        c.add(ng.CustomNoiseItem(1.0, noise_type=t))
        return c

    blended = ng.CustomNoiseChain()
    blended.add(ng.BlendedNoise(1.0, normalize=None, blend_function=sb.hostutil.BLENDING_MODES["lerp"],
                                custom_noise_1=chain_of("pyramid"), custom_noise_2=chain_of("perlin"), noise_2_percent=0.5))
    sched = ng.CustomNoiseChain()
    sched.add(ng.ScheduledNoise(1.0, noise=blended, start_sigma=10.0, end_sigma=1.0, normalize=None, fallback_noise=chain_of("gaussian")))
    x = torch.zeros(16, 16, 128, 128, device=dev)
    ns = sched.make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0)
    s, sn_ = torch.tensor(5.0), torch.tensor(4.5)

    def c3():
        torch.manual_seed(0)
        return ns(s, sn_)

    us, per = _time_traced(sb, c3, 5, flush)
    record("C3 Scheduled(Blended(pyramid, perlin)) 16x16x128x128", x.numel(), 16.8, us, per,
           "16.8 B/el = injected-draw accounting of SURVEY 8d; the run also writes its own Philox draws")

    # C4: wavelet CFG db2 / 3 levels / separate H,V,D scales on 16x4x128x128
    class _MS:
        sigma_min, sigma_max = torch.tensor(0.03), torch.tensor(14.6)

        @staticmethod
        def timestep(sg):
            return (sg.log() - math.log(0.03)) / (math.log(14.6) - math.log(0.03)) * 999

    class _Model:
        model_sampling = _MS()

    cond, uncond, xin = (torch.randn(16, 4, 128, 128, device=dev) for _ in range(3))
    fn = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(
        wave="db2", level=3, diff={"yl_scale": 5, "yh_scales": [[3, 4, 5]] * 3}))
    wargs = {"sigma": torch.full((16,), 5.0, device=dev), "input": xin, "cond_denoised": cond, "uncond_denoised": uncond,
             "cond_scale": 7.0, "model": _Model(), "model_options": {}}  # fmt: skip
    us, per = _time_traced(sb, lambda: fn(wargs), 10, flush)
    record("C4 wavelet CFG db2 L3 16x4x128x128 (fp64 coefficients)", xin.numel(), 16.0, us, per,
           "16 B/el: read cond, uncond, x; write result (fp64 is internal)")

    # C5 per-GPU shard: video latent 1x16x33x90x160, power noise via frames_to_channels + DPM++ SDE half steps
    x5 = torch.zeros(1, 16, 33, 90, 160, device=dev)
    params = ng.CustomNoiseParametersNoise(
        1.0, noise=power_chain(), normalize=None, override_device=None, override_dtype=None, frames_to_channels=True,
        ensure_square_aspect_ratio=False, fix_invalid=False, rng_mode="default", rng_offset_mode="disabled", rng_state_offset=0)
    c5 = ng.CustomNoiseChain()
    c5.add(params)
    ns5 = c5.make_noise_sampler(x5, None, None, seed=0)
    us, per = _time_traced(sb, lambda: ns5(None, None), 5, flush)
    record("C5 shard power noise 1x16x33x90x160", x5.numel(), 8.125 + 8.125 + 12, us, per, "as C1")

    sig5 = torch.tensor([14.6, 7.0, 2.0, 0.7], device=dev)
    xv = torch.randn(1, 16, 33, 90, 160, device=dev) * 14.6
    cfg = {"noise_type": "gaussian"}

    def c5_dpm():
        torch.manual_seed(0)
        return sb.samplers.SonarDPMPPSDE.sampler(lambda x, s, **k: x * 0.9, xv, sig5, extra_args={"seed": 0}, disable=True, sonar_params=cfg)

    us, per = _time_traced(sb, c5_dpm, 3, flush)
    record("C5 shard sonar_dpmpp_sde 3 steps 1x16x33x90x160 (fused Gaussian noise)", 3 * xv.numel(), 40.0, us, per,
           "40 B/el/step: two fused half steps x 20 B/el; Philox moments pre-passes cost no HBM bytes")
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE.json configs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
