"""ncu target: perlin-only and pyramid-only fused samples."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import sonar_b200 as sb
from test_gpu_fused_noise import chain_of
dev = torch.device("cuda", 0)
x = torch.zeros(16, 16, 128, 128, device=dev)
for kind in ("perlin", "pyramid"):
    ns = chain_of(sb, kind).make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0)
    for _ in range(3):
        out = ns(None, None)
torch.cuda.synchronize()
print(float(out.std()))
