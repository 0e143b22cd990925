import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easy_neural_ode_b200 import mnist_step
dev = torch.device("cuda:0"); B = 100
def batch(s, pin=False):
    g = torch.Generator().manual_seed(1234 + s)
    t = (torch.randint(0, 256, (B, 28, 28, 1), generator=g, dtype=torch.uint8), torch.randint(0, 10, (B,), generator=g),
         torch.randint(0, 2, (B, 784), generator=g).float() * 2 - 1)
    return tuple(x.pin_memory() for x in t) if pin else tuple(x.to(dev) for x in t)
bs = [batch(s) for s in range(8)]; hs = [batch(s, True) for s in range(8)]
m = mnist_step.MnistNode(device=dev)
snap = m.snapshot()
orig_capture = m._capture
caps = []
def cap(*a):
    t0 = time.perf_counter(); orig_capture(*a); torch.cuda.synchronize(); caps.append((time.perf_counter() - t0) * 1e3)
m._capture = cap
m._budget = None
for phase in ("device", "host"):
    m.restore(snap); ts = []
    for s in range(30):
        torch.cuda.synchronize(); t0 = time.perf_counter(); n0 = len(caps)
        if phase == "device": o = m.update_graphed(*bs[s % 8])
        else:
            h = hs[s % 8]; o = m.update_graphed(*(x.to(dev, non_blocking=True) for x in h))
        torch.cuda.synchronize(); ts.append(((time.perf_counter() - t0) * 1e3, len(caps) - n0, o["redone"], o["fwd_iters"], o["bwd_iters"], sorted(getattr(m, "_graphs", {}).keys())))
    print(phase, "total ms", sum(t[0] for t in ts))
    for i, t in enumerate(ts): print("  ", i, f"{t[0]:.2f}ms", "captures", t[1], "redone", t[2], "iters", t[3], t[4], "budget", t[5])
print("capture times", [round(c, 1) for c in caps])
