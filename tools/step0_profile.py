"""cProfile of the first sampler step's host path (look-ahead planning + general fused step): a 2-step run."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb  # noqa: E402

dev = torch.device("cuda", 0)
sig = torch.tensor([14.6, 7.0, 3.0, 0.0], device=dev)
x0 = torch.randn(8, 4, 128, 128, device=dev) * 14.6
sampler = sb.samplers.SonarEulerAncestral


class M:
    def __init__(self):
        self.t = []
        self.last = None

    def __call__(self, x, sigma, **_kw):
        now = time.perf_counter()
        if self.last is not None:
            self.t.append(now - self.last)
        d = x * 0.9
        self.last = time.perf_counter()
        return d


def run(m):
    return sampler.sampler(m, x0, sig, extra_args={"seed": 0}, disable=True)


for _ in range(5):
    run(M())
torch.cuda.synchronize()
acc = []
for _ in range(50):
    m = M()
    run(m)
    acc.append(m.t)
torch.cuda.synchronize()
for i in range(2):
    vals = sorted(a[i] for a in acc)
    print(f"host time after model call {i}: median {vals[len(vals) // 2] * 1e6:.1f} us")
prof = cProfile.Profile()
prof.enable()
for _ in range(100):
    run(M())
torch.cuda.synchronize()
prof.disable()
pstats.Stats(prof).sort_stats("cumtime").print_stats(45)
