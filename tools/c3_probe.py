"""C3 noise graph: per-kernel times of the fused and the materialising paths (CUDA events around every launch)."""
import statistics, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from test_gpu_fused_noise import c3_graph, chain_of

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
x = torch.zeros(16, 16, 128, 128, device=dev)
s, sn = torch.tensor(5.0), torch.tensor(4.5)
for name, make in (("c3", lambda: c3_graph(sb)), ("pyramid", lambda: chain_of(sb, "pyramid")), ("perlin", lambda: chain_of(sb, "perlin"))):
    for fused in (True, False):
        sb.generators.FUSED_NOISE = fused
        sb.generators.FUSED_SINGLE_GENERATORS = fused
        ns = make().make_noise_sampler(x, torch.tensor(0.03), torch.tensor(14.6), seed=0)
        for _ in range(5):
            ns(s, sn)
        per, tot = {}, []
        for _ in range(10):
            flush.zero_()
            sb.ops.TRACE = []
            ns(s, sn)
            torch.cuda.synchronize()
            tr, sb.ops.TRACE = sb.ops.TRACE, None
            t = 0.0
            for n, a, b in tr:
                us = a.elapsed_time(b) * 1e3
                per.setdefault(n, []).append(us); t += us
            tot.append(t)
        # device time of the whole sample with the host running ahead (8 x 256 MiB of zero-fill in front of it)
        dev_us = []
        for _ in range(10):
            for _ in range(8):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ns(s, sn); e1.record(); torch.cuda.synchronize()
            dev_us.append(e0.elapsed_time(e1) * 1e3)
        print(f"{name:8s} fused={fused!s:5s} device {statistics.median(dev_us):7.1f} us | traced", end=" ")
        print(f"{statistics.median(tot):7.1f} us  " + ", ".join(f"{k.replace('sonar_','')}:{statistics.median(v):.1f}" for k, v in per.items()))
