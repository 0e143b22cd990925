import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import mnist_step
dev = torch.device("cuda:0")
B, D = 100, 784
gen = torch.Generator().manual_seed(0)
images = torch.randint(0, 256, (B, 28, 28, 1), generator=gen, dtype=torch.uint8).to(dev)
labels = torch.randint(0, 10, (B,), generator=gen).to(dev)
eps = (torch.randint(0, 2, (B, D), generator=gen).float() * 2 - 1).to(dev)
m = mnist_step.MnistNode(reg="none", device=dev, reg_type="fin", lam_fro=1e-2, lam_kin=1e-2)
for _ in range(2):
    m.update(images, labels, eps)
spec = eno.MLPDynamics(D, 100)
p = spec.init_params(seed=0, device=dev)
x0 = images.float().reshape(B, -1) / 255.
for _ in range(2):
    eno.odeint_vmap(spec.dynamics_wrap, x0, torch.tensor([0., 1.]), p)
torch.cuda.synchronize()
print("ok")
