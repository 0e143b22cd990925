import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import ode as eode
from oracle import dynamics_oracle as do, ode_oracle as oo
cuda = torch.device("cuda:0")
Bn, D, H = 32, 784, 100
spec = eno.MLPDynamics(D, H, reg="r3")
p_cpu = do.init_mlp_params(D, H, seed=0, scale=1.5)
gen = torch.Generator().manual_seed(2)
x0 = torch.rand((Bn, D), generator=gen)
eps = torch.randint(0, 2, (Bn, D), generator=gen).float() * 2 - 1
ts = torch.tensor([0., 1.])
cl = do.make_mnist_closures("r3")
z = torch.zeros(Bn)
tr = []
want, nfe_w, st = oo.odeint(cl["all_aug_dynamics"], (x0, z, z, z), ts, eps, p_cpu, rtol=1.4e-8, atol=1.4e-8, return_stats=True, trace=tr)
p_gpu = {k: {kk: vv.to(cuda) for kk, vv in v.items()} for k, v in p_cpu.items()}
zg = z.to(cuda)
got, nfe_g = eode._odeint_augmented(spec.all_aug_dynamics, (x0.to(cuda), zg, zg, zg), ts, (eps.to(cuda), p_gpu), 1.4e-8, 1.4e-8, float("inf"), trace_cap=64)
trg = eno.last_stats["fwd_trace"]
for i in range(max(len(tr), len(trg))):
    print(i, "O", tr[i] if i < len(tr) else None, "C", trg[i].tolist() if i < len(trg) else None)
for a, b, n in zip(got, want, "y r fro kin".split()):
    print(n, float((a[-1].cpu() - b[-1]).abs().max() / b[-1].abs().max()))
