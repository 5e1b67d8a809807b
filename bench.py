"""Benchmark of the sonar_b200 hot path (driver contract: ONE JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], "C5", the configuration north_star shards over 8 GPUs and the largest one that
fits a single GPU): video latent 8x16x33x90x160 (Hunyuan-style 5-D), `sonar_dpmpp_sde` whose custom noise is
SonarCustomNoiseParameters(frames_to_channels=True) o SonarPowerNoise(alpha=1) -- exactly the reference objects of
py/noise.py:2103-2185 + py/nodes/powernoise.py:355-408 + py/sonar.py:649-770 -- for N_SAMPLER_STEPS sampler steps
(each: 2 model evaluations, 2 power-noise samples, 2 fused momentum half steps; the last step is the sigma_next == 0
Euler step). Denoiser stub x*0.9 outside the timed regions. One bench "step" = one such sampling run over the whole
8-item batch. Metric: latent noise elements / second, elements = N_SAMPLER_STEPS x 60,825,600 per run
(the unit of work is the latent, per sampler step).

* value     : device-resident throughput. Everything the sampler enqueues between two model calls (power-noise
              sample + fused half step) is bracketed by CUDA events on the launching stream; the stub denoiser
              in between zero-fills FLUSH_PASSES x 256 MiB (L2 eviction, and enough device time that the host
              enqueues the next half step while the device is still "in the DiT": the intervals hold device
              time, not Python launch latency). ms_per_step = sum of those intervals, max over ranks.
* e2e       : the same run through the public sampler function with HOST buffers: pinned x0 (this rank's shard)
              -> H2D, the run, NCCL gather of the shards to rank 0 (N > 1), D2H of the whole result into pinned
              host memory; wall clock between device synchronisations, max over ranks.
* roofline  : the kernel with the largest share of the timed region, from a traced run (a CUDA-event pair around
              every C-ABI launch on the launching stream): algorithmic bytes per launch / its mean
              duration, against MEASURED_PEAKS.json hbm_gbs; `kernels` lists every kernel's share.
* cpu_baseline : the CPU oracle port of the reference algorithm (oracle/sonar_oracle.py) on the host cores, on a
              bounded sample of the same workload (one batch item, a few sampler steps).
* N > 1     : STRONG scaling: the 8-item batch is split over the ranks (8 GPUs: one video latent each, as north_star
              states); every rank regenerates its slice of the global Philox draws; the two doubles of each
              scale_noise travel through the NVLink peer mailboxes (no data-path collective); `parity_max_abs_diff`
              is the gathered result vs an un-sharded run of the whole batch on rank 0.
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

# SONAR_BENCH_ITEMS=k (diagnostics only): run the job on k batch items, e.g. 1 = the per-GPU work of the 8-GPU split
SHAPE = (int(os.environ.get("SONAR_BENCH_ITEMS", "8")), 16, 33, 90, 160)
ITEM_ELEMS = SHAPE[1] * SHAPE[2] * SHAPE[3] * SHAPE[4]
N_SAMPLER_STEPS = 10
FLUSH_PASSES = 16  # 256 MiB zero-fills per stub-denoiser call: L2 eviction + device time for the host to run ahead
ELEMS_PER_RUN = N_SAMPLER_STEPS * SHAPE[0] * ITEM_ELEMS
WORKLOAD = (
    "C5 video latent 8x16x33x90x160: sonar_dpmpp_sde with frames_to_channels power noise (alpha=1) as custom noise, "
    f"{N_SAMPLER_STEPS} sampler steps"
)
# SURVEY.md 8d, RNG-fused accounting: a power-noise sample writes 4 B/el (nothing is read); a fused half step reads
# x, denoised, history, noise and writes x', history' = 24 B/el
BYTES_NOISE, BYTES_STEP = 4.0, 24.0


def make_sigmas(n: int = N_SAMPLER_STEPS) -> torch.Tensor:
    return torch.cat((torch.linspace(14.6, 0.03, n), torch.zeros(1)))


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
            # nvidia-smi spends ~0.1-1 s initialising NVML: wait for its first sample so that start-up does not land
            # inside the first timed runs; the 100 ms polling that follows is what samples the timed region
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
def cpu_reference_run(n_runs: int, sampler_steps: int = 2, items: int = SHAPE[0]) -> dict:
    """Times the oracle port of the C5 job on the CPU: oracle.power_noise (irfft2 of the shaped complex draw,
    py/nodes/powernoise.py:355-366) + scale_noise + SonarOracle.dpmpp_sde (py/sonar.py:649-735), all host threads,
    on `items` batch items and `sampler_steps` steps (all of them two-stage steps: the schedule does not reach 0)."""
    from oracle import sonar_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sigmas = make_sigmas()[: sampler_steps + 1]
    shape = (items, *SHAPE[1:])
    folded = (items, SHAPE[1] * SHAPE[2], SHAPE[3], SHAPE[4])
    filt = orc.power_filter(folded, alpha=1.0)
    torch.manual_seed(0)
    x0 = torch.randn(shape) * sigmas[0]

    def model(x, _sigma):
        return x * 0.9

    def noise():
        spec = torch.randn((*folded[:-1], folded[-1] // 2 + 1), dtype=torch.complex64)
        raw = orc.power_noise(iter([spec]), folded, filt, normalized=False)
        return orc.scale_noise(raw.reshape(shape), 1.0, normalized=True)

    times = []
    for _ in range(n_runs):
        o = orc.SonarOracle()
        x = x0.clone()
        t_run = 0.0
        for i in range(sampler_steps):
            den = model(x, sigmas[i])  # untimed, like the GPU arm's stub
            t0 = time.perf_counter()
            n1, n2 = noise(), noise()
            x = o.dpmpp_sde(i, x, den, sigmas[i], sigmas[i + 1], model, n1, n2)
            t_run += time.perf_counter() - t0
        times.append(t_run)
    best = min(times)
    elems = sampler_steps * x0.numel()
    per_el = best / elems
    return {
        "value": 1.0 / per_el,
        "unit": "elements/s",
        "cores": cores,
        "kind": "port",
        "sample": f"best of {n_runs} x {sampler_steps} sampler steps of C5 on {items} of 8 batch items (oracle port on CPU; the "
        "second model call of each step is inside the timed region, x*0.9)",
        "ms_per_step": per_el * ELEMS_PER_RUN * 1e3,
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    for _ in range(min(1, args.warmup)):
        cpu_reference_run(1, sampler_steps=1, items=2)
    res = cpu_reference_run(max(1, min(args.steps, 3)))
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
        "scaling": "strong",
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
    """Denoiser stub that doubles as the boundary of the timed regions: everything the sampler enqueues between
    two model calls is the hot path of one half step."""

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
            # the GPUs meet HERE, with the host already enqueueing the denoiser and the first half step behind the
            # rendezvous: from its release on every rank runs from a full queue
            self.on_first_call()
            self.on_first_call = None
        if self.flush is not None:
            for _ in range(self.flush_passes):
                self.flush.zero_()
        den = x * 0.9
        if self.timed:
            self._open = torch.cuda.Event(enable_timing=True)
            self._open.record()
        return den

    def total_ms(self) -> float:
        return sum(a.elapsed_time(b) for a, b in self.pairs)


POWER_KW = dict(time_brownian=False, alpha=1.0, max_freq=0.7071, min_freq=0.0, stretch=1.0, rotate=0.0, pnorm=2.0,
                mix=1.0, common_mode=0.0, channel_correlation="1, 1, 1, 1, 1, 1")  # fmt: skip


def power_chain(sb, **kw):
    c = sb.noise_graph.CustomNoiseChain()
    c.add(sb.spectral_noise.PowerNoiseItem(1.0, **(POWER_KW | kw)))
    return c


def c5_chain(sb):
    """SonarCustomNoiseParameters(frames_to_channels=True) o SonarPowerNoise(alpha=1)."""
    item = sb.noise_graph.CustomNoiseParametersNoise(
        1.0, noise=power_chain(sb), normalize=None, override_device=None, override_dtype=None, frames_to_channels=True,
        ensure_square_aspect_ratio=False, fix_invalid=False, rng_mode="default", rng_offset_mode="disabled", rng_state_offset=0,
    )  # fmt: skip
    chain = sb.noise_graph.CustomNoiseChain()
    chain.add(item)
    return chain


def sampler_run(sb, model, x0, sigmas, chain):
    return sb.samplers.SonarDPMPPSDE.sampler(
        model, x0, sigmas, extra_args={"seed": 0}, disable=True, sonar_params={"custom_noise": chain},
    )


# which launches of the C5 run are which kernel, and their algorithmic bytes per element of the tensor they produce
C5_KERNELS = {
    "sonar_spectral_filter_f32": ("spectral_batched_kernel (power-noise sample: gain -> irfft2 -> moments; its complex Philox draw is sonar_philox_*)", BYTES_NOISE),
    "sonar_step_f32": ("sonar_step_fast_vec_kernel (fused half step, noise normalised on load)", BYTES_STEP),
}


def traced_breakdown(sb, run, n_local_elems: int, peak: float, peak_src: str, reps: int) -> tuple[dict, list]:
    """Per-kernel CUDA-event timing of the C5 run (ops.TRACE brackets every C-ABI launch on the launching stream).
    Returns the roofline block of the kernel with the largest share and the table of all kernels. Launches may batch
    several units (look-ahead: several noise samples per FFT launch), so bytes are counted per run and divided by
    the launches of the run."""
    per: dict[str, list[float]] = {}
    for _ in range(reps):
        sb.ops.TRACE = []
        run()
        torch.cuda.synchronize()
        trace, sb.ops.TRACE = sb.ops.TRACE, None
        for name, a, b in trace:
            per.setdefault(name, []).append(a.elapsed_time(b) * 1e3)
    total = sum(sum(v) for v in per.values())
    units = {"sonar_spectral_filter_f32": 2 * (N_SAMPLER_STEPS - 1), "sonar_step_f32": 2 * (N_SAMPLER_STEPS - 1) + 1}
    table = []
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        label, bytes_per_el = C5_KERNELS.get(name, (name, None))
        launches = len(v) // reps
        row = {"entry_point": name, "kernel": label, "launches_per_run": launches, "mean_us": statistics.mean(v),
               "us_per_run": sum(v) / reps, "share_of_timed_region": sum(v) / total}  # fmt: skip
        if bytes_per_el is not None:
            algo = bytes_per_el * n_local_elems * units[name] / launches  # per launch
            row |= {"units_per_run": units[name], "algorithmic_bytes_per_launch": algo,
                    "achieved_gbs": algo / (statistics.mean(v) * 1e-6) / 1e9, "frac": algo / (statistics.mean(v) * 1e-6) / 1e9 / peak}  # fmt: skip
        table.append(row)
    top = next(r for r in table if "frac" in r)
    roofline = {
        "bound": "hbm",
        "kernel": top["kernel"],
        "achieved": top["achieved_gbs"],
        "peak": peak,
        "unit": "GB/s",
        "frac": top["frac"],
        "traffic": None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"],
        "launch_us": top["mean_us"],
        "launches_timed": top["launches_per_run"] * reps,
        "share_of_timed_region": top["share_of_timed_region"],
        "timing": "CUDA-event pair around every launch of the timed C5 run on the launching stream (tensors of 243 MB per "
        "GPU at N=1 exceed L2; the stub denoiser flushes L2 between half steps)",
    }
    return roofline, table


def run_b200_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; sonar_b200 has no CPU path (use --impl reference for the CPU oracle)")
    if world > SHAPE[0]:
        raise SystemExit(f"bench.py: the C5 batch has {SHAPE[0]} items; at most {SHAPE[0]} ranks")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT when the first communicator is created; stdout carries the ONE
        # JSON line, so the communicator is created here with file descriptor 1 pointing at stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    import sonar_b200 as sb

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sigmas_host = make_sigmas()
    sigmas = sigmas_host.to(dev)
    sizes = sb.parallel.split_sizes(SHAPE[0], world)
    b0, nb = sum(sizes[:rank]), sizes[rank]
    # identical synthetic latent on every rank (one CPU generator), then this rank's batch slice
    gen = torch.Generator().manual_seed(1234)
    x0_host = (torch.randn(SHAPE, generator=gen) * sigmas_host[0])[b0 : b0 + nb].contiguous().pin_memory()
    x0 = x0_host.to(dev)
    n_local = x0.numel()
    chain = c5_chain(sb)

    def one_run(model, src, sharded=True):
        torch.manual_seed(99)  # replicated generator: every rank reserves the same global draws
        if world > 1 and sharded:
            with sb.parallel.sharded(SHAPE[0], rank=rank, world_size=world):
                return sampler_run(sb, model, src, sigmas, chain)
        return sampler_run(sb, model, src, sigmas, chain)

    def device_rendezvous():
        with sb.parallel.sharded(SHAPE[0], rank=rank, world_size=world):
            sb.parallel.device_barrier()

    rendezvous = device_rendezvous if world > 1 else None

    # ---------------- device-resident throughput ----------------
    for _ in range(max(3, args.warmup)):
        barrier()
        one_run(StepTimer(dev, timed=False, on_first_call=rendezvous), x0)
    barrier()
    launches0 = sb.ops.LAUNCH_COUNT
    timers = []
    gc.collect()
    gc.disable()  # a collection pause between two launches would show up as device idle time
    with ClockSampler(local_rank, enabled=rank == 0 and os.environ.get("SONAR_BENCH_NO_CLOCKS") is None) as clocks:
        barrier()
        one_run(StepTimer(dev, timed=False, on_first_call=rendezvous), x0)  # one more untimed run with the clock poller up
        barrier()
        launches0 = sb.ops.LAUNCH_COUNT
        t_wall = time.perf_counter()
        for _ in range(args.steps):
            if world > 1:
                barrier()
            timer = StepTimer(dev, on_first_call=rendezvous)
            t_run = time.perf_counter()
            one_run(timer, x0)
            timer.close()
            timer.host_ms = (time.perf_counter() - t_run) * 1e3  # host time to ENQUEUE the run (not a device time)
            timers.append(timer)
        barrier()
        wall = time.perf_counter() - t_wall
        launches = sb.ops.LAUNCH_COUNT - launches0
        # keep the sampler busy long enough for a few clock samples (a FIXED run count: sharded runs exchange)
        keep = StepTimer(dev, timed=False)
        for _ in range(10):
            one_run(keep, x0)
        torch.cuda.synchronize()
    gc.enable()
    run_ms = [t.total_ms() for t in timers]
    ms_local = statistics.mean(run_ms)
    per_rank = torch.tensor([ms_local], device=dev, dtype=torch.float64)
    if world > 1:
        gathered = [torch.zeros_like(per_rank) for _ in range(world)]
        dist.all_gather(gathered, per_rank)
        per_rank_ms = [round(float(g.item()), 4) for g in gathered]
    else:
        per_rank_ms = [round(ms_local, 4)]
    ms_per_step = max(per_rank_ms)
    value = ELEMS_PER_RUN / (ms_per_step * 1e-3)

    # ---------------- roofline of the dominant kernel (per-launch CUDA events in the same run) ----------------
    peak, peak_src = measured_peak()
    barrier()
    tracer = StepTimer(dev, timed=False, on_first_call=rendezvous)
    # per-kernel durations are taken with the noise pipeline OFF: with it on, the producers run beside the step kernel
    # on a second stream and an event pair around one launch also times its neighbours (ncu serialises in the same way)
    pipelined = bool(sb.samplers.NOISE_PIPELINE and n_local >= (sb.samplers.PIPELINE_MIN_NUMEL_SHARDED if world > 1 else sb.samplers.PIPELINE_MIN_NUMEL))
    sb.samplers.NOISE_PIPELINE, keep_flag = False, sb.samplers.NOISE_PIPELINE
    try:
        one_run(tracer, x0)
        roofline, kernel_table = traced_breakdown(sb, lambda: one_run(tracer, x0), n_local, peak, peak_src, reps=2)
    finally:
        sb.samplers.NOISE_PIPELINE = keep_flag
    roofline["timing"] += (
        "; this traced run has the noise pipeline switched off (kernels one after the other, as under ncu) -- in the timed "
        "runs the producers of the next sample overlap the step kernel, see roofline_job"
    ) if pipelined else ""
    # the job as a whole: SURVEY 8d bytes of every half step (24 B/el) and every noise sample (4 B/el) over the timed region
    job_bytes = (BYTES_STEP * (2 * (N_SAMPLER_STEPS - 1) + 1) + BYTES_NOISE * 2 * (N_SAMPLER_STEPS - 1)) * n_local
    roofline_job = {
        "bound": "hbm",
        "algorithmic_bytes_per_run": job_bytes,
        "achieved": job_bytes / (ms_local * 1e-3) / 1e9,
        "peak": peak,
        "unit": "GB/s",
        "frac": job_bytes / (ms_local * 1e-3) / 1e9 / peak,
        "noise_pipeline": pipelined,
        "note": "this rank's algorithmic bytes of the whole run (19 half steps x 24 B/el + 18 noise samples x 4 B/el) / its "
        "device time; the noise producers are instruction-bound (Philox + Box-Muller, 90x160 FFT), which is what keeps this "
        "below the step kernel's own fraction",
    }
    traffic_path = REPO / "profiles" / "traffic.json"
    if traffic_path.exists():
        roofline["traffic"] = json.loads(traffic_path.read_text()).get("c5:" + roofline["kernel"].split(" ")[0])

    # ---------------- multi-GPU parity: gathered shards vs the un-sharded job on rank 0 ----------------
    parity = None
    if world > 1:
        barrier()
        out = one_run(StepTimer(dev, flush_bytes=0, timed=False), x0)
        with sb.parallel.sharded(SHAPE[0], rank=rank, world_size=world):
            full = sb.parallel.gather(out, dst=0)
        if rank == 0:
            x_full = (torch.randn(SHAPE, generator=torch.Generator().manual_seed(1234)) * sigmas_host[0]).to(dev)
            want = one_run(StepTimer(dev, flush_bytes=0, timed=False), x_full, sharded=False)
            parity = float((full - want).abs().max().item())
            del x_full, want
        del full, out
        barrier()

    # ---------------- end to end through the public API with host buffers ----------------
    e2e_times = []
    e2e_model = StepTimer(dev, flush_bytes=0, timed=False)
    gc.collect()
    gc.disable()
    out_host = torch.empty(SHAPE if rank == 0 else (nb, *SHAPE[1:]), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    for i in range(2 + args.steps):
        barrier()
        t0 = time.perf_counter()
        x_dev = x0_host.to(dev, non_blocking=True)
        out = one_run(e2e_model, x_dev)
        if world > 1:
            if rank == 0:
                # this rank's own shard goes to the host while NCCL gathers the others over NVLink
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    out_host[b0 : b0 + nb].copy_(out, non_blocking=True)
            with sb.parallel.sharded(SHAPE[0], rank=rank, world_size=world):
                full = sb.parallel.gather(out, dst=0)
            if rank == 0:
                out_host[nb:].copy_(full[nb:], non_blocking=True)
        else:
            out_host.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
        if i >= 2:
            e2e_times.append(time.perf_counter() - t0)
    gc.enable()
    # where the end-to-end time goes (one extra pass, events on the stream; not part of the timed numbers)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    t0 = time.perf_counter()
    ev[0].record()
    x_dev = x0_host.to(dev, non_blocking=True)
    ev[1].record()
    out = one_run(e2e_model, x_dev)
    ev[2].record()
    t_enq = time.perf_counter() - t0
    if world == 1:
        out_host.copy_(out, non_blocking=True)
    ev[3].record()
    torch.cuda.synchronize()
    e2e_breakdown = {"h2d_ms": ev[0].elapsed_time(ev[1]), "run_ms": ev[1].elapsed_time(ev[2]), "host_enqueue_ms": t_enq * 1e3}
    if world == 1:
        e2e_breakdown["d2h_ms"] = ev[2].elapsed_time(ev[3])
    e2e_s = statistics.mean(e2e_times)
    t_dev = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    e2e_s = float(t_dev.item())
    bytes_total = SHAPE[0] * ITEM_ELEMS * 4
    e2e = {
        "value": ELEMS_PER_RUN / e2e_s,
        "unit": "elements/s",
        "h2d_bytes_per_step": bytes_total,
        "d2h_bytes_per_step": bytes_total,
        "ms_per_step": e2e_s * 1e3,
        "breakdown_rank0": {k: round(v, 3) for k, v in e2e_breakdown.items()},
        "note": "public sampler function (SonarDPMPPSDE.sampler, custom_noise chain), pinned host x0 -> H2D (each rank its "
        "shard), the sampler run with the stub denoiser inside, "
        + ("NCCL gather to rank 0 overlapped with rank 0's own D2H, " if world > 1 else "")
        + "whole result D2H into pinned memory; bytes are job totals",
    }

    extras = None
    cpu_baseline = None
    if rank == 0 and world == 1:
        del x0, out
        torch.cuda.empty_cache()
        if not args.no_extras:
            extras = run_extras(sb, dev, peak)
        cpu = cpu_reference_run(2, items=4)
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
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "global_batch": SHAPE[0],
                "global_shape": list(SHAPE),
                "items_per_gpu": sizes,
                "sampler_steps": N_SAMPLER_STEPS,
                "parallelism": f"batch-sharded x{world} (strong split)" if world > 1 else "single GPU",
                "l2": f"tensors of {bytes_total // world >> 20} MiB per GPU; {FLUSH_PASSES} x 256 MiB zero-fill between half steps "
                "(stub denoiser, untimed) evicts L2",
                "timing": "sum of the CUDA-event intervals between model calls (power-noise sample + fused half step), max over ranks",
                "noise_pipeline": "on: the next sample is produced on a second stream beside the step kernel, joined before the "
                "next model call" if pipelined else "off (samples per GPU below PIPELINE_MIN_NUMEL: batched look-ahead)",
                "elements": "sampler steps x latent elements (each step = 2 noise samples + 2 fused half steps)",
            },
            "roofline": roofline,
            "roofline_job": roofline_job,
            "kernels": kernel_table,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "wall_s_timed_region": wall,
            "runs_ms": run_ms,
            "runs_host_enqueue_ms": [round(t.host_ms, 2) for t in timers],
            "runs_max_interval_ms": [round(max(a.elapsed_time(b) for a, b in t.pairs), 3) for t in timers],
            "per_rank_ms": per_rank_ms,
        }
        if parity is not None:
            line["parity_max_abs_diff"] = parity
            line["parity_note"] = "max |gathered sharded result - un-sharded run of the whole batch on rank 0|, same seed"
        if extras is not None:
            line["other_configs"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# the other BASELINE.json configs (single GPU): elements/s and roofline fraction, SURVEY 8d byte accounting
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
        # ~0.3 ms of queued zero-fills (they also evict L2): the host enqueues the traced launches while the device is
        # still busy, so an event pair brackets the kernel alone and not the host's launch latency on an idle GPU
        for _ in range(8):
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
                "launches": len(per),
                "elements_per_s": elems / (total_us * 1e-6),
                "algorithmic_bytes_per_element": algo_bytes_per_elem,
                "hbm_gbs": elems * algo_bytes_per_elem / (total_us * 1e-6) / 1e9,
                "frac_of_measured_peak": elems * algo_bytes_per_elem / (total_us * 1e-6) / 1e9 / peak,
                "top_kernel": top,
                "per_kernel_us": per,
                "note": note,
            },
        )

    ng = sb.noise_graph
    # C1: SonarPowerNoise pink on 1x4x64x64
    x = torch.zeros(1, 4, 64, 64, device=dev)
    ns = power_chain(sb).make_noise_sampler(x, None, None, seed=0)
    us, per = _time_traced(sb, lambda: ns(None, None), 10, flush)
    record("C1 SonarPowerNoise pink 1x4x64x64", x.numel(), 4.0, us, per,
           "SURVEY 8d RNG-fused accounting: 4 B/el (the sample is written once; the chain's own normalisation pass is overhead)")

    # C2: sonar_euler_ancestral, fused Gaussian noise, 8x4x128x128, 30 steps (round-1 headline workload)
    sig2 = torch.cat((torch.linspace(14.6, 0.03, 30), torch.zeros(1))).to(dev)
    x2 = torch.randn(8, 4, 128, 128, device=dev) * 14.6

    def c2_model(x, _s, **_k):
        flush.zero_()  # the UNet's place: evicts L2 and lets the host run ahead of the device (see _time_traced)
        return x * 0.9

    def c2():
        torch.manual_seed(99)
        return sb.samplers.SonarEulerAncestral.sampler(c2_model, x2, sig2, extra_args={"seed": 0}, disable=True)

    us, per = _time_traced(sb, c2, 5, flush)
    record("C2 sonar_euler_ancestral 8x4x128x128, 30 steps, fused Gaussian noise", 30 * x2.numel(), 20.0, us, per,
           "20 B/el/step: read x, denoised, history; write x', history'; noise regenerated from Philox in registers. Kernel "
           "time only, cold L2 (a 256 MiB zero-fill stands in for the UNet between steps); latency-bound at this size: see DESIGN.md")

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
    ns3 = sched.make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0)
    s, sn_ = torch.tensor(5.0), torch.tensor(4.5)

    def c3():
        torch.manual_seed(0)
        return ns3(s, sn_)

    us, per = _time_traced(sb, c3, 5, flush)
    record("C3 Scheduled(Blended(pyramid, perlin)) 16x16x128x128", x.numel(), 4.0, us, per,
           "SURVEY 8d RNG-fused accounting: 4 B/el (one write of the result); 16.8 B/el with injected base draws")

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
