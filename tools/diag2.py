import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno
from oracle import dynamics_oracle as do, ode_oracle as oo
from tests.util import load, params_from_flat, rel_err
cuda = torch.device("cuda:0")
g = load("mlp_forward.npz")
B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
spec = eno.MLPDynamics(D, H)
pc = params_from_flat(g["params"], D, H)
pg = {k: {kk: vv.to(cuda) for kk, vv in v.items()} for k, v in pc.items()}
x0 = torch.as_tensor(g["x0"])
ts = torch.tensor([0., 0.3])
tr = []
ys_o, nfe_o, st = oo.odeint(lambda y, t, p: do.mlp_dynamics(p, y, t), x0, ts, pc, rtol=1e-7, atol=1e-7, mxstep=2, return_stats=True, trace=tr)
print("oracle", st, tr)
ys_g, stg, trg = eno.solve_with_trace(spec, "dopri5", x0.to(cuda), ts, pg, 1e-7, 1e-7, mxstep=2)
print("cuda", stg, trg)
print("err vs golden", rel_err(ys_g, g["fwd_mx2_y"]), "oracle vs golden", rel_err(ys_o, g["fwd_mx2_y"]))
print(ys_g[1,0,:6], ys_o[1,0,:6])
# converged answer
ys_c, _ = oo.odeint(lambda y, t, p: do.mlp_dynamics(p, y, t), x0.double(), ts.double(), {k: {kk: vv.double() for kk, vv in v.items()} for k, v in pc.items()}, rtol=1e-10, atol=1e-10)
print("cuda vs converged", rel_err(ys_g, ys_c), "oracle vs converged", rel_err(ys_o, ys_c))
