"""Fused step with in-register Gaussian noise (SONAR_NOISE_PHILOX_NORMALIZED) at video-latent sizes: per-launch time
from CUDA events around the launches of a short sonar_dpmpp_sde run (default Gaussian noise, look-ahead statistics)."""
import statistics, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6439.5
for shape in ((1, 16, 33, 90, 160), (8, 16, 33, 90, 160), (16, 16, 128, 128)):
    sig = torch.tensor([14.6, 9.0, 5.0, 2.0, 0.7], device=dev)
    x = torch.randn(shape, device=dev) * 14.6

    def model(v, s, **k):
        flush.zero_()
        return v * 0.9

    def run():
        torch.manual_seed(0)
        return sb.samplers.SonarDPMPPSDE.sampler(model, x, sig, extra_args={"seed": 0}, disable=True, sonar_params={"noise_type": "gaussian"})

    for _ in range(3):
        run()
    per = {}
    for _ in range(5):
        sb.ops.TRACE = []
        run()
        torch.cuda.synchronize()
        tr, sb.ops.TRACE = sb.ops.TRACE, None
        for n, a, b in tr:
            per.setdefault(n, []).append(a.elapsed_time(b) * 1e3)
    n = x.numel()
    us = statistics.median(per["sonar_step_f32"])
    print(f"{shape}: sonar_step_f32 (Philox noise in registers) {us:7.1f} us = {20 * n / us / 1e3:7.1f} GB/s = {20 * n / us / 1e3 / peak:.3f} of measured peak;"
          f" others: " + ", ".join(f"{k.replace('sonar_', '')}:{statistics.median(v):.1f}" for k, v in per.items() if k != "sonar_step_f32"))
