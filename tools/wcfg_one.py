"""ncu target: the C4 wavelet-CFG call."""
import math, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb
dev = torch.device("cuda", 0)
class _MS:
    sigma_min, sigma_max = torch.tensor(0.03), torch.tensor(14.6)
    @staticmethod
    def timestep(sg):
        return (sg.log() - math.log(0.03)) / (math.log(14.6) - math.log(0.03)) * 999
class _Model:
    model_sampling = _MS()
cfg = sb.wcfg.WaveletCFG(existing_cfg=None, rules=sb.wcfg.WCFGRules.build(wave="db2", level=3, diff={"yl_scale": 5, "yh_scales": [[3, 4, 5]] * 3}))
cond, uncond, xin = (torch.randn(16, 4, 128, 128, device=dev) for _ in range(3))
wargs = {"sigma": torch.full((16,), 5.0, device=dev), "input": xin, "cond_denoised": cond, "uncond_denoised": uncond, "cond_scale": 7.0, "model": _Model(), "model_options": {}}
for _ in range(4):
    out = cfg(wargs)
torch.cuda.synchronize()
print(float(out.std()))
