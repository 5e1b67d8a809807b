"""Round-2 probe: does the ALU-bound power-noise sample (Philox fill + FFT) overlap with an HBM-bound streaming kernel
(stand-in for the fused step: 24 B/el) when they run on two streams? Prints serial vs concurrent time per items."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb

dev = torch.device("cuda", 0)
H, W, Wh = 90, 160, 81
mask = torch.rand(H, Wh, device=dev) + 0.5
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
side = torch.cuda.Stream(device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for items in (1, 2, 8):
    spec = torch.empty(items, 528, H, Wh, dtype=torch.complex64, device=dev)
    out = torch.empty(items, 528, H, W, device=dev)
    n_floats = spec.numel() * 2
    std = 0.5 ** 0.5
    a, b, c, d = (torch.randn(items, 528, H, W, device=dev) for _ in range(4))
    o1, o2 = torch.empty_like(a), torch.empty_like(a)

    def noise():
        torch.manual_seed(1)
        dr = sb.ops.reserve_draw(n_floats, dev)
        sb.ops.philox_fill(dr, spec, kind="normal", p0=0.0, p1=std)
        sb.ops.spectral_filter(spectrum=spec, mask=mask, hw=(H, W), out_scale=1.0 / 120.0, out=out)

    def stream_step():  # 4 reads + 2 writes of the latent = 24 B/el
        torch.add(a, b, out=o1)
        torch.add(c, d, out=o2)

    def serial():
        noise(); stream_step()

    def concurrent(noise_first=True):
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        if noise_first:
            with torch.cuda.stream(side):
                noise()
            stream_step()
        else:
            stream_step()
            with torch.cuda.stream(side):
                noise()
        main.wait_stream(side)

    tn, ts_, tser = timeit(noise), timeit(stream_step), timeit(serial)
    tc1, tc2 = timeit(lambda: concurrent(True)), timeit(lambda: concurrent(False))
    print(f"items={items}: noise {tn:7.1f}  step-proxy {ts_:7.1f}  serial {tser:7.1f}  concurrent(noise first) {tc1:7.1f}  (step first) {tc2:7.1f} us")
