"""Host-side cost of one sampler step (the e2e number is host-bound at config C2's tensor sizes).

    python tools/host_profile.py [runs]      -> wall time per step + cProfile top functions
"""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb  # noqa: E402

dev = torch.device("cuda", 0)
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 50
sig = torch.cat((torch.linspace(14.6, 0.03, 30), torch.zeros(1))).to(dev)
x0 = torch.randn(8, 4, 128, 128, device=dev) * 14.6


def model(x, sigma, **_kw):
    return x * 0.9


def run():
    return sb.samplers.SonarEulerAncestral.sampler(model, x0, sig, extra_args={"seed": 0}, disable=True)


for _ in range(5):
    run()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(runs):
    run()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / runs
print(f"sampler run: {dt * 1e6:.1f} us  ({dt * 1e6 / 30:.2f} us per step, host + device, stub model inside)")

# the floor: 30 x (stub model + empty_like) without our step
t0 = time.perf_counter()
for _ in range(runs):
    x = x0
    for i in range(30):
        d = model(x, sig[i] * 1.0)
        x = torch.empty_like(x)
torch.cuda.synchronize()
dt0 = (time.perf_counter() - t0) / runs
print(f"floor (30 x stub model, sigma index, empty_like): {dt0 * 1e6:.1f} us ({dt0 * 1e6 / 30:.2f} us per step)")

prof = cProfile.Profile()
prof.enable()
for _ in range(runs):
    run()
torch.cuda.synchronize()
prof.disable()
st = pstats.Stats(prof)
st.sort_stats("tottime").print_stats(28)
