"""Development check: graphed update() == polled update() (same trajectory), and its speed."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easy_neural_ode_b200 import mnist_step
dev = torch.device("cuda:0"); B = 100
def batch(s):
    g = torch.Generator().manual_seed(1234 + s)
    return (torch.randint(0, 256, (B, 28, 28, 1), generator=g, dtype=torch.uint8).to(dev),
            torch.randint(0, 10, (B,), generator=g).to(dev),
            (torch.randint(0, 2, (B, 784), generator=g).float() * 2 - 1).to(dev))
bs = [batch(s) for s in range(8)]
a = mnist_step.MnistNode(device=dev); b = mnist_step.MnistNode(device=dev)
K = 12
la = [float(a.update(*bs[s % 8])["loss"]) for s in range(K)]
lb = []
for s in range(K):
    o = b.update_graphed(*bs[s % 8]); lb.append(o["loss"])
    print(s, "polled", la[s], "graphed", o)
print("max |loss diff|", max(abs(x - y) for x, y in zip(la, lb)), "param diff", float((a.p_ode - b.p_ode).abs().max()))
snap = b.snapshot()
for name, fn in (("graphed", b.update_graphed), ("polled", a.update)):
    m = b if name == "graphed" else a
    m.restore(snap) if name == "graphed" else None
    ts = []
    for s in range(20):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        o = fn(*bs[s % 8])
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "ms per step:", " ".join(f"{t:.2f}" for t in ts), "| mean", sum(ts) / len(ts))
