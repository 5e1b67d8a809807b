"""Host-side cost of the C5 job (SONAR_BENCH_ITEMS items): wall time to ENQUEUE one run and cProfile of the host path."""
import cProfile, pstats, sys, time
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


def model(x, sigma, **_kw):
    return x * 0.9


def run():
    torch.manual_seed(99)
    return bench.sampler_run(sb, model, x0, sig, chain)


for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ts.append(((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
print(f"items={bench.SHAPE[0]}: host enqueue {min(t[0] for t in ts):.2f} ms per run, enqueue + drain {min(t[1] for t in ts):.2f} ms")
prof = cProfile.Profile()
prof.enable()
for _ in range(10):
    run()
torch.cuda.synchronize()
prof.disable()
st = pstats.Stats(prof)
st.sort_stats("tottime").print_stats(30)
st.sort_stats("cumulative").print_stats(25)
