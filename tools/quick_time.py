"""Quick device timing of the C1 train step (not the bench contract; a development probe)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import mnist_step

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1.4e-8
gen = torch.Generator().manual_seed(0)
images = torch.randint(0, 256, (B, 28, 28, 1), generator=gen, dtype=torch.uint8).to(dev)
labels = torch.randint(0, 10, (B,), generator=gen).to(dev)
eps = (torch.randint(0, 2, (B, 784), generator=gen).float() * 2 - 1).to(dev)
m = mnist_step.MnistNode(reg="r3", lam=6e-5, rtol=tol, atol=tol, device=dev)
for _ in range(3):
    out = m.loss_and_grads(images, labels, eps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    out = m.loss_and_grads(images, labels, eps)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"B={B} tol={tol} train-step {ms:.3f} ms  fwd {out['fwd_stats']}  bwd {out['bwd_stats']}")
# forward only
x0 = (images.float() / 255.).reshape(B, -1)
for _ in range(3):
    eno.odeint(m.spec.dynamics_wrap, x0, m.ts, m.p_ode, rtol=tol, atol=tol)
e0.record()
for _ in range(n):
    eno.odeint(m.spec.dynamics_wrap, x0, m.ts, m.p_ode, rtol=tol, atol=tol)
e1.record(); torch.cuda.synchronize()
print(f"forward solve {e0.elapsed_time(e1)/n:.3f} ms")
# per-kernel device times from the library's profiling hooks
import ctypes as C
from easy_neural_ode_b200 import _cabi
L = _cabi.lib(); h = _cabi.ctx(0)
L.eno_profile_enable(h, 1)
for _ in range(5):
    m.loss_and_grads(images, labels, eps)
torch.cuda.synchronize()
L.eno_profile_enable(h, 0)
ms = (C.c_double * 5)(); cnt = (C.c_int64 * 5)()
L.eno_profile_read(h, ms, cnt)
for name, a, b in zip(["fwd_step", "adj_rows", "adj_pgrad"], ms, cnt):
    print(f"{name}: {1e3*a/max(b,1):.1f} us avg over {b} launches")
