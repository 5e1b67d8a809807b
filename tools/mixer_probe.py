"""Channel-mixer GEMM at the C5 shard size: bulk-async pipelined kernel vs the plain tiled kernel (forced by a
4-byte-misaligned input) vs torch.matmul (cuBLAS SGEMM) on the reference's swapped layout."""
import sys, statistics
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


for shape in ((1, 528, 90, 160), (8, 528, 90, 160), (16, 16, 128, 128), (8, 64, 64, 64)):
    b, c, h, w = shape
    torch.manual_seed(0)
    mixer = torch.randn(c, c, device=dev) / c ** 0.5
    buf = torch.randn(b * c * h * w + 4, device=dev)
    aligned = buf[:-4].view(shape)
    misaligned = buf[1:-3].view(shape)
    flops = 2.0 * b * c * c * h * w
    packed = sb.ops.pack_mixer(mixer)
    t_bulk = timeit(lambda: sb.ops.channel_mix(aligned, mixer, None, packed))
    out_mis = torch.empty(b * c * h * w + 4, device=dev)[1:-3].view(shape)
    t_tiled = timeit(lambda: sb.ops.channel_mix(aligned, mixer))
    t_torch = timeit(lambda: (mixer @ aligned.swapaxes(0, 1).reshape(c, -1)).reshape(c, b, h, w).swapaxes(1, 0).contiguous())
    ref = (mixer.double() @ aligned.double().swapaxes(0, 1).reshape(c, -1)).reshape(c, b, h, w).swapaxes(1, 0)
    err = (sb.ops.channel_mix(aligned, mixer, None, packed).double() - ref).abs().max().item()
    print(f"{shape}: bulk-async {t_bulk:8.1f} us ({flops / t_bulk / 1e6:6.1f} TFLOP/s)  tiled {t_tiled:8.1f} us  torch (swap + cuBLAS + swap) {t_torch:8.1f} us  max err {err:.2e}")
