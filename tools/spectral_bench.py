"""Times sonar_spectral_filter_f32 alone (CUDA events, L2 flushed between launches) over a few plane shapes,
spectrum and real input -- the yardstick used while tuning csrc/spectral.cu.

    python tools/spectral_bench.py [reps]
"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb  # noqa: E402

dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PEAK = 6439.5e9

CASES = [  # (planes, H, W)
    (528, 90, 160),   # C5 shard: 1 x 16 x 33 frames
    (4, 64, 64),      # C1
    (32, 128, 128),   # SDXL batch 8
    (256, 128, 128),  # Flux batch 16
    (2560, 32, 32),   # FreeU stage-1 activations, batch 2
    (1280, 64, 64),   # FreeU stage-2 activations, batch 2
    (64, 256, 256),   # generic kernel (spectrum does not fit shared memory)
]


def timed(fn):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for planes, h, w in CASES:
    wh = w // 2 + 1
    spec = torch.randn(planes, h, wh, dtype=torch.complex64, device=dev)
    real = torch.randn(planes, h, w, device=dev)
    mask = torch.rand(h, wh, device=dev) + 0.5
    f_spec = lambda: sb.ops.spectral_filter(spectrum=spec, mask=mask, hw=(h, w), out_scale=1 / math.sqrt(h * w))  # noqa: E731
    f_real = lambda: sb.ops.spectral_filter(real=real, mask=mask, hw=(h, w), out_scale=1 / (h * w))  # noqa: E731
    for name, fn, bytes_el in (("spectrum", f_spec, 8.0 * wh / w + 4.0), ("real", f_real, 8.0)):
        for _ in range(3):
            fn()
        med, best = timed(fn)
        n = planes * h * w
        print(f"{planes:5d} x {h:3d} x {w:3d} {name:8s}: median {med:8.1f} us  best {best:8.1f} us  "
              f"{n / med * 1e-3:7.2f} G el/s  {n * bytes_el / (med * 1e-6) / PEAK:5.3f} of HBM peak", flush=True)
