"""ncu target: one Philox normal fill of the C5 half spectrum (61.6 M floats)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb
dev = torch.device("cuda", 0)
spec = torch.empty(8, 528, 90, 81, dtype=torch.complex64, device=dev)
for _ in range(3):
    torch.manual_seed(1)
    d = sb.ops.reserve_draw(spec.numel() * 2, dev)
    sb.ops.philox_fill(d, spec, kind="normal", p0=0.0, p1=0.5 ** 0.5)
torch.cuda.synchronize()
print(float(spec.real.std()))
