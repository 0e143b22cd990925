import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from easy_neural_ode_b200 import mnist_step
dev = torch.device("cuda:0"); B = 100
mode = sys.argv[1] if len(sys.argv) > 1 else "random"
gp = torch.Generator().manual_seed(99)
proj = torch.randn((784, 10), generator=gp)
def batch(s):
    g = torch.Generator().manual_seed(1234 + s)
    im = torch.randint(0, 256, (B, 28, 28, 1), generator=g, dtype=torch.uint8)
    if mode == "random":
        lab = torch.randint(0, 10, (B,), generator=g)
    else:
        lab = ((im.reshape(B, -1).float() / 255. - 0.5) @ proj).argmax(dim=1)
    eps = torch.randint(0, 2, (B, 784), generator=g).float() * 2 - 1
    return im.to(dev), lab.to(dev), eps.to(dev)
m = mnist_step.MnistNode(device=dev)
for s in range(40):
    o = m.update(*batch(s % 8 if mode != "fresh" else s))
    if s % 3 == 0 or s > 35:
        print(s, round(float(o["loss"]), 4), o["fwd_stats"]["n_iters"], o["bwd_stats"][0]["n_iters"])
