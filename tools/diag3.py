import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno
from oracle import dynamics_oracle as do, ode_oracle as oo
from tests.util import load, params_from_flat, rel_err
cuda = torch.device("cuda:0")
g = load("mlp_adjoint.npz")
B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
spec = eno.MLPDynamics(D, H, reg="r3")
params = params_from_flat(g["params"], D, H, device=cuda)
x0 = torch.as_tensor(g["x0"], device=cuda); eps = torch.as_tensor(g["eps"], device=cuda)
z = torch.zeros(B, device=cuda)
(ys, r, fro, kin), nfe = eno.odeint(spec.all_aug_dynamics, (x0, z, z, z), torch.tensor([0., 1.]), eps, params, rtol=1e-5, atol=1e-5)
print("nfe", float(nfe), float(g["all_nfe"]), eno.last_stats["fwd"])
print("y", rel_err(ys[-1], g["all_y1"]), "r", rel_err(r[-1], g["all_r"]), "fro", rel_err(fro[-1], g["all_fro"]), "kin", rel_err(kin[-1], g["all_kin"]))
print(r[-1].cpu(), g["all_r"]); print(fro[-1].cpu(), g["all_fro"]); print(kin[-1].cpu(), g["all_kin"])
# single rhs check
dy, rr = eno.rhs(spec, 1, x0, 0.3, params)
cl = do.make_mnist_closures("r3")
pc = {k: {kk: vv.cpu() for kk, vv in v.items()} for k, v in params.items()}
w = cl["all_aug_dynamics"]((x0.cpu(),), 0.3, eps.cpu(), pc)
print("rhs r", rel_err(rr, w[1]))
