"""Development: per-phase cycle breakdown of the cluster adjoint kernel (needs libeno_phases.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ENO_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "easy-neural-ode_b200", "libeno_phases.so")
import ctypes as C, torch
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import mnist_step, _cabi
dev = torch.device("cuda:0"); B = 100
gen = torch.Generator().manual_seed(0)
images = torch.randint(0, 256, (B, 28, 28, 1), generator=gen, dtype=torch.uint8).to(dev)
labels = torch.randint(0, 10, (B,), generator=gen).to(dev)
eps = (torch.randint(0, 2, (B, 784), generator=gen).float() * 2 - 1).to(dev)
m = mnist_step.MnistNode(reg="r3", lam=6e-5, device=dev)
L = _cabi.lib()
L.eno_debug_phases.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
cyc = (C.c_longlong * 32)(); cnt = (C.c_longlong * 32)()
for _ in range(2): m.loss_and_grads(images, labels, eps)
L.eno_debug_phases(cyc, cnt)
x0 = (images.float() / 255.).reshape(B, -1)
# only the ADJOINT step kernel starts/flushes the clocks (forward-kernel ticks land in a dead slot array)
out = m.loss_and_grads(images, labels, eps)
n = L.eno_debug_phases(cyc, cnt)
names = ["prologue", "stage input(+sync)", "mv over Dl (thread 0)", "mv over H (thread 0)", "allreduce: end sync", "reduce_local body+sync", "store(+elem+sync)", "scalar", "combine", "wait mv stragglers", "allred local", "allred cluster.sync", "allred remote+epi"]
tot = sum(cyc[i] for i in range(n))
print("one train step (fwd+bwd kernels, CTA 0), total cycles", tot)
for i in range(n):
    print(f"{names[i]:14s} cycles {cyc[i]:9d} {100*cyc[i]/tot:5.1f}%  count {cnt[i]:6d}  avg {cyc[i]/max(cnt[i],1):8.1f}")
