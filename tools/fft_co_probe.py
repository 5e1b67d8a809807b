"""FFT kernel alone, default form vs the co-scheduled low-register form (sonar_set_grid_limit(3)), 4224 planes of 90x160."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb
dev = torch.device("cuda", 0)
H, W, Wh = 90, 160, 81
mask = torch.rand(H, Wh, device=dev) + 0.5
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
import os
P = int(os.environ.get("PLANES", "4224"))
spec = torch.randn(P, H, Wh, dtype=torch.complex64, device=dev)
out = torch.empty(P, H, W, device=dev)
ref = None
for limit in (0, 3, 2):
    ts = []
    for i in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sb.ops.set_grid_limit(limit)
        a.record()
        sb.ops.spectral_filter(spectrum=spec, mask=mask, hw=(H, W), out_scale=1.0 / 120.0, out=out)
        b.record()
        sb.ops.set_grid_limit(0)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    if ref is None:
        ref = out.clone()
    print(f"planes {P} limit {limit}: {sorted(ts)[len(ts) // 2]:7.1f} us   max |diff vs default| {float((out - ref).abs().max()):.2e}")
