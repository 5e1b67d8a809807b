"""ncu target: a few C3 samples through the fused path."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import sonar_b200 as sb
from test_gpu_fused_noise import c3_graph
dev = torch.device("cuda", 0)
x = torch.zeros(16, 16, 128, 128, device=dev)
ns = c3_graph(sb).make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0)
for _ in range(3):
    out = ns(torch.tensor(5.0), torch.tensor(4.5))
torch.cuda.synchronize()
print(float(out.std()))
