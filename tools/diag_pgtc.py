"""Development probe: the adjoint of a small plain-dynamics problem on the cluster path (tensor-core parameter-gradient
contraction) and on the streamed path (FFMA contraction), controller traces of the first segment next to the oracle's.
Usage: [ENO_PATH=stream] python tools/diag_pgtc.py [B]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import ode
from oracle import dynamics_oracle as do, ode_oracle as oo
from tests.util import flat_from_params

Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 6
D, H = 16, 8
dev = torch.device("cuda", 0)
p_cpu = do.init_mlp_params(D, H, seed=4, scale=2.0)
gen = torch.Generator().manual_seed(11)
x0 = torch.rand((Bn, D), generator=gen)
gys = torch.randn((5, Bn, D), generator=gen)
ts = torch.tensor([0., 0.2, 0.5, 0.6, 1.0])
for tol in (1e-6, 1e-5):
    f = lambda y, t, p: do.mlp_dynamics(p, y.reshape(Bn, D), t).reshape(-1)
    ys_w, nfe_w, st = oo.solve(lambda y, t: f(y, t, p_cpu), x0.reshape(-1), ts, tol, tol)
    bst = []
    want = oo.odeint_rev(f, tol, tol, math.inf, ys_w, ts, (p_cpu,), gys.reshape(5, -1), bst)
    print("tol", tol, "oracle first segment:", [(round(t, 5), round(dt, 6), float("%.4g" % r), a) for t, dt, r, a in bst[0]["trace"]])
    spec = eno.MLPDynamics(D, H)
    pflat = flat_from_params(p_cpu).to(dev)
    for path in (os.environ.get("ENO_PATH", "default"),):   # the context reads ENO_PATH once: run the script per path
        ys, stf, _ = ode._fwd_call(spec, "dopri5", x0.to(dev), ts.to(dev), pflat, tol, tol, math.inf)
        out = ode._bwd_call(spec, 0, False, ys, ts.to(dev), pflat, gys.to(dev), torch.zeros((5, Bn), device=dev), tol, tol, math.inf,
                            trace_cap=64)
        tr = out[5][: out[4][0]["n_iters"]].cpu().tolist()
        print("  ", path, [s["n_iters"] for s in out[4]], [(round(t, 5), round(dt, 6), float("%.4g" % r), bool(a)) for t, dt, r, a in tr])
