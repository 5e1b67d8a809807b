"""ncu target: the spectral kernel on `planes` planes of 90x160 (spectrum input). usage: spectral_one.py [planes] [reps]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb

planes = int(sys.argv[1]) if len(sys.argv) > 1 else 528
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
H, W = 90, 160
spec = torch.randn(planes, H, W // 2 + 1, dtype=torch.complex64, device=dev)
mask = torch.rand(H, W // 2 + 1, device=dev) + 0.5
out = torch.empty(planes, H, W, device=dev)
for _ in range(reps):
    sb.ops.spectral_filter(spectrum=spec, mask=mask, hw=(H, W), out_scale=1.0 / 120.0, out=out)
torch.cuda.synchronize()
print("ok", float(out.std()))
