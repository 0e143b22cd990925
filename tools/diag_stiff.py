"""Development (GPU box): controller trajectories (t, dt, ratio, accepted) of the CUDA path and the oracle on the stiff golden
case (tests/golden/mlp_stiff.npz, rejected steps on both solves), forward and adjoint."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import ode, _cabi
from oracle import dynamics_oracle as do, ode_oracle as oo
from tests.util import load, params_from_flat, flat_from_params, rel_err

cuda = torch.device("cuda:0")
g = load("mlp_stiff.npz")
B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
tol = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-5
params = params_from_flat(g["params"], D, H)
x0, eps, gy = (torch.as_tensor(g[k]) for k in ("x0", "eps", "gy"))
cl = do.make_mnist_closures("r3")
ts = torch.tensor([0., 1.])
trf = []
ys_o, nfe_o, st = oo.odeint(cl["fwd"], x0, ts, eps, params, rtol=tol, atol=tol, return_stats=True, trace=trf)
res = (ys_o, torch.zeros(2, B, 1))
gg = (torch.zeros_like(res[0]), torch.zeros_like(res[1])); gg[0][-1] = gy; gg[1][-1] = 0.01
bst = []
out_o = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, params), gg, bst)
spec = eno.MLPDynamics(D, H, reg="r3")
pflat = torch.as_tensor(g["params"], device=cuda)
ys, stf, trg = ode._fwd_call(spec, "dopri5", x0.to(cuda), ts.to(cuda), pflat, tol, tol, math.inf, trace_cap=256)
print("forward: oracle", st, "cuda", stf, " y1 rel", rel_err(ys[-1], ys_o[-1]))
for i, o in enumerate(trf):
    print(i, "O", tuple(round(v, 9) if isinstance(v, float) else v for v in o), "C", [round(v, 9) for v in trg[i].tolist()])
g_y = torch.zeros((2, B, D), device=cuda); g_y[1] = gy.to(cuda)
g_r = torch.zeros((2, B), device=cuda); g_r[1] = 0.01
# feed the adjoint the ORACLE's forward trajectory so that both adjoint solves start from identical data
outc = ode._bwd_call(spec, _cabi.AUG_REG, True, ys_o.to(cuda).contiguous(), ts.to(cuda), pflat, g_y, g_r, tol, tol, math.inf, trace_cap=256)
print("adjoint (same ys): oracle", {k: v for k, v in bst[0].items() if k != "trace"}, "cuda", outc[4][0])
tro, trc = bst[0]["trace"], outc[5][: outc[4][0]["n_iters"]].cpu()
for i in range(max(len(tro), trc.shape[0])):
    o = tuple(round(v, 9) if isinstance(v, float) else v for v in tro[i]) if i < len(tro) else None
    c = [round(v, 9) for v in trc[i].tolist()] if i < trc.shape[0] else None
    print(i, "O", o, "C", c)
print("x0_bar rel", rel_err(outc[0], out_o[0][0]), "params_bar rel", rel_err(outc[3], flat_from_params(out_o[3])))
