"""Development probe: controller traces (t, dt, ratio, accepted) of a tableau solve on the CUDA path next to the oracle's.
Usage: python tools/diag_sweep.py <tableau> <tol> [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easy_neural_ode_b200 as eno
from oracle import dynamics_oracle as do, ode_oracle as oo

name, tol = sys.argv[1], float(sys.argv[2])
Bn = int(sys.argv[3]) if len(sys.argv) > 3 else 100
D, H = 784, 100
dev = torch.device("cuda", 0)
spec = eno.MLPDynamics(D, H)
p_cpu = do.init_mlp_params(D, H, seed=0, scale=2.0)
x0 = torch.rand((Bn, D), generator=torch.Generator().manual_seed(3))
ts = torch.tensor([0., 1.])
p_gpu = {k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p_cpu.items()}
fun = lambda y, t: do.mlp_dynamics(p_cpu, y.reshape(Bn, D), t).reshape(-1)
tr_o = []
want, nfe_w, st = oo.solve(fun, x0.reshape(-1), ts, tol, tol, solver=name, trace=tr_o)
fun64 = lambda y, t: do.mlp_dynamics({k: {kk: vv.double() for kk, vv in v.items()} for k, v in p_cpu.items()}, y.reshape(Bn, D), t).reshape(-1)
tr_64 = []
oo.solve(fun64, x0.double().reshape(-1), ts.double(), tol, tol, solver=name, trace=tr_64)
got, stg, tr_g = eno.solve_with_trace(spec, name, x0.to(dev), ts, p_gpu, tol, tol, trace_cap=16384)
print(name, tol, "nfe oracle", float(nfe_w), "cuda", stg["nfe"])
for i in range(max(len(tr_o), len(tr_g), len(tr_64))):
    f = lambda tr: ("t %.7f dt %.6g ratio %.6g %s" % (tr[i][0], tr[i][1], tr[i][2], bool(tr[i][3]))) if i < len(tr) else "-"
    print(i, "| oracle fp32:", f(tr_o), "| cuda:", f(tr_g.tolist()), "| oracle fp64:", f(tr_64))
