import sys, time, statistics
sys.path.insert(0, "/root/repo")
import torch, bench
import sonar_b200 as sb
dev = torch.device("cuda", 0)
sig = bench.make_sigmas().to(dev)
x0 = torch.randn(bench.SHAPE, device=dev) * 14.6
def run(model):
    torch.manual_seed(99)
    return bench.sampler_run(sb, model, x0, sig)
warm = bench.StepTimer(dev, timed=False)
for _ in range(5): run(warm)
torch.cuda.synchronize()
firsts=[]; rests=[]; host0=[]
for _ in range(8):
    torch.cuda.synchronize()
    t = bench.StepTimer(dev)
    run(t); t.close(); torch.cuda.synchronize()
    iv=[a.elapsed_time(b)*1e3 for a,b in t.pairs]
    firsts.append(iv[0]); rests.append(sum(iv[1:])/len(iv[1:]))
print("step0 us", [round(v,1) for v in firsts]); print("other steps avg us", [round(v,2) for v in rests])
# host time of the first step's path: time from model return to next model call, measured on host
import cProfile, pstats
class HostTimer:
    def __init__(self): self.t=[]; self.last=None
    def __call__(self, x, sigma, **kw):
        now=time.perf_counter()
        if self.last is not None: self.t.append(now-self.last)
        d = x*0.9
        self.last=time.perf_counter()
        return d
for _ in range(3):
    h=HostTimer(); run(h); torch.cuda.synchronize()
    print("host us: step0 %.1f, others avg %.1f" % (h.t[0]*1e6, statistics.mean(h.t[1:])*1e6))
