"""Device time of the C5 job (bench.py's timed region only) for the current SONAR_B200_* knobs; SONAR_BENCH_ITEMS=k
runs k of the 8 batch items (1 = the per-GPU share of the 8-GPU split)."""
import statistics, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
import sonar_b200 as sb

dev = torch.device("cuda", 0)
sig_h = bench.make_sigmas()
sig = sig_h.to(dev)
x0 = (torch.randn(bench.SHAPE, generator=torch.Generator().manual_seed(1234)) * sig_h[0]).to(dev)
chain = bench.c5_chain(sb)


def run(model):
    torch.manual_seed(99)
    return bench.sampler_run(sb, model, x0, sig, chain)


for _ in range(3):
    out = run(bench.StepTimer(dev, timed=False))
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t = bench.StepTimer(dev)
    out = run(t)
    t.close()
    torch.cuda.synchronize()
    ts.append(t.total_ms())
iv = [a.elapsed_time(b) * 1e3 for a, b in t.pairs]
print(f"items={bench.SHAPE[0]} pipeline={sb.samplers.NOISE_PIPELINE} step_ctas={sb.samplers.PIPELINE_STEP_CTAS} fill_ctas={sb.samplers.PIPELINE_FILL_CTAS} fft_ctas={sb.samplers.PIPELINE_FFT_CTAS} split={sb.samplers.PIPELINE_STEP_SPLIT} stepB={sb.samplers.PIPELINE_STEP_B_CTAS} "
      f"chunk={sb.samplers.NOISE_PIPELINE_CHUNK}: {statistics.median(ts):.3f} ms per run  intervals(us): " + " ".join(f"{v:.0f}" for v in iv)
      + f"  checksum {float(out.double().sum()):.6f}")
