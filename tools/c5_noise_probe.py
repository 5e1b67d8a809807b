"""Round-2 probe: how should the C5 power-noise sample (complex Philox draw -> gain -> irfft2) be scheduled?
  A  one fill launch + one spectral launch over the whole tensor (spectrum round-trips HBM)
  B  per-item chunks through ONE reusable spectrum buffer (stays in L2)
  C  B with fill(i+1) on a second stream, overlapped with spectral(i)
Also: the spectral kernel alone at 528 / 1056 / 4224 planes (tail effect of the persistent grid)."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb

dev = torch.device("cuda", 0)
H, W, Wh = 90, 160, 81
mask = (torch.rand(H, Wh, device=dev) + 0.5)
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
    ts.sort()
    return ts[len(ts) // 2]


for items in (1, 8):
    planes = items * 528
    spec = torch.empty(items, 528, H, Wh, dtype=torch.complex64, device=dev)
    out = torch.empty(items, 528, H, W, device=dev)
    chunk = torch.empty(2, 528, H, Wh, dtype=torch.complex64, device=dev)
    n_floats = spec.numel() * 2
    per_item = n_floats // items
    std = 0.5 ** 0.5

    def draw():
        torch.manual_seed(1)
        return sb.ops.reserve_draw(n_floats, dev)

    def A():
        d = draw()
        sb.ops.philox_fill(d, spec, kind="normal", p0=0.0, p1=std)
        sb.ops.spectral_filter(spectrum=spec, mask=mask, hw=(H, W), out_scale=1.0 / 120.0, out=out)

    def fill_only():
        sb.ops.philox_fill(draw(), spec, kind="normal", p0=0.0, p1=std)

    def spectral_only():
        sb.ops.spectral_filter(spectrum=spec, mask=mask, hw=(H, W), out_scale=1.0 / 120.0, out=out)

    def B():
        d = draw()
        for i in range(items):
            buf = chunk[i & 1]
            sb.ops.philox_fill(d, buf, kind="normal", p0=0.0, p1=std, begin=i * per_item)
            sb.ops.spectral_filter(spectrum=buf, mask=mask, hw=(H, W), out_scale=1.0 / 120.0, out=out[i])

    side = torch.cuda.Stream(device=dev)

    def C():
        d = draw()
        main = torch.cuda.current_stream()
        evs = []
        side.wait_stream(main)
        for i in range(items):
            with torch.cuda.stream(side):
                sb.ops.philox_fill(d, chunk[i & 1], kind="normal", p0=0.0, p1=std, begin=i * per_item)
                e = torch.cuda.Event(); e.record(side); evs.append(e)
            main.wait_event(evs[i])
            sb.ops.spectral_filter(spectrum=chunk[i & 1], mask=mask, hw=(H, W), out_scale=1.0 / 120.0, out=out[i])
            if i + 1 < items:
                e2 = torch.cuda.Event(); e2.record(main)
                # buffer (i+1)&1 was last read by spectral(i-1), already ordered before spectral(i) on main
                if i >= 1:
                    side.wait_event(e_prev)
                e_prev = e2
        main.wait_stream(side)

    A(); ref = out.clone(); B(); torch.cuda.synchronize()
    print(f"items={items}: B==A {torch.equal(ref, out)}", end="  ")
    C(); torch.cuda.synchronize()
    print(f"C==A {torch.equal(ref, out)}")
    print(f"  A total {timeit(A):8.1f} us   fill {timeit(fill_only):8.1f}   spectral {timeit(spectral_only):8.1f}")
    print(f"  B chunked {timeit(B):8.1f} us   C two-stream {timeit(C):8.1f} us")

for planes in (148, 296, 444, 528, 592, 1056, 2112, 4224):
    spec = torch.randn(planes, H, Wh, dtype=torch.complex64, device=dev)
    out = torch.empty(planes, H, W, device=dev)
    t = timeit(lambda: sb.ops.spectral_filter(spectrum=spec, mask=mask, hw=(H, W), out_scale=1.0 / 120.0, out=out))
    print(f"spectral {planes:5d} planes: {t:8.1f} us  ({t / planes * 528:6.1f} us per 528)")
