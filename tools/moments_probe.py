import sys, torch
sys.path.insert(0, "/root/repo")
import sonar_b200 as sb
from sonar_b200 import ops
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
n = 524288
draw = ops.reserve_draw(n, dev)
offsets = [draw.offset + j * draw.counter_offset for j in range(29)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(): return ops.philox_normal_moments_batch(draw, offsets, begin=0, count=n, device=dev)
for _ in range(3): run()
ts=[]
for _ in range(10):
    flush.zero_()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s = run(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b)*1e3)
print(sorted(ts)[len(ts)//2], s[0].tolist())
