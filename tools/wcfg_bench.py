"""Times the wavelet-CFG call (config C4 rule: db2, 3 levels, fp64 coefficients) for a few batch sizes, CUDA events
with an L2 flush between calls."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


class _MS:
    sigma_min, sigma_max = torch.tensor(0.03), torch.tensor(14.6)

    @staticmethod
    def timestep(sg):
        return (sg.log() - math.log(0.03)) / (math.log(14.6) - math.log(0.03)) * 999


class _Model:
    model_sampling = _MS()


cfg = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(
    wave="db2", level=3, diff={"yl_scale": 5, "yh_scales": [[3, 4, 5]] * 3}))
for batch in (2, 8, 16, 18, 19, 37, 64, 256):
    cond, uncond, xin = (torch.randn(batch, 4, 128, 128, device=dev) for _ in range(3))
    wargs = {"sigma": torch.full((batch,), 5.0, device=dev), "input": xin, "cond_denoised": cond, "uncond_denoised": uncond,
             "cond_scale": 7.0, "model": _Model(), "model_options": {}}
    for _ in range(3):
        cfg(wargs)
    ts, ks = [], []
    for _ in range(12):
        flush.zero_()
        torch.cuda.synchronize()
        n0 = sb.ops.LAUNCH_COUNT
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sb.ops.TRACE = []
        a.record()
        cfg(wargs)
        b.record()
        b.synchronize()
        trace, sb.ops.TRACE = sb.ops.TRACE, None
        ts.append(a.elapsed_time(b) * 1e3)
        ks.append(sum(s.elapsed_time(e) for _, s, e in trace) * 1e3)
    ts.sort()
    ks.sort()
    n = batch * 4 * 128 * 128
    print(f"batch {batch:3d} ({batch * 4:3d} planes): kernel median {ks[len(ks) // 2]:7.1f} us, best {ks[0]:7.1f} us "
          f"({16 * n / ks[len(ks) // 2] * 1e-3:7.1f} GB/s at 16 B/el); call median {ts[len(ts) // 2]:7.1f} us, launches {sb.ops.LAUNCH_COUNT - n0}")
