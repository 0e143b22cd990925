"""Development diagnostic for the FFJORD tabular path: GPU vs the fp32 oracle vs an fp64 oracle solve at 1e-10 ("truth"),
at several tolerances.  Prints iteration counts and max-norm relative errors of y(t1) and the gradients, so that the
parity tolerances in tests/test_gpu_ffjord.py are measured, not guessed.   python tools/diag_ffjord.py [B D H]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easy_neural_ode_b200 as eno
from oracle import ffjord_oracle as fo, ode_oracle as oo


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def main():
    B, D, H = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 43, 128)
    dev = torch.device("cuda:0")
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    p = fo.init_css_params(D, H, seed=3, scale=1.3)
    g = torch.Generator().manual_seed(9)
    x = torch.randn((B, D), generator=g)
    eps = torch.randn((B, D), generator=g)
    gy = torch.randn((B, D), generator=torch.Generator().manual_seed(4))
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    spec = eno.ConcatSquashDynamics(D, H, reg="r2")
    cl = fo.make_ffjord_closures("r2", "our")

    def oracle(dtype, tol):
        pp = {k: {kk: vv.to(dtype) for kk, vv in v.items()} for k, v in p.items()}
        zz = z.to(dtype)
        res, nfe, fst = oo.odeint_split_fwd(cl["fwd"], (x.to(dtype), zz, zz), ts.to(dtype), eps.to(dtype), pp, n_aux=2,
                                            rtol=tol, atol=tol, return_stats=True)
        gg = tuple(torch.zeros_like(r) for r in res)
        gg[0][-1], gg[1][-1], gg[2][-1] = gy.to(dtype), 1.0 / B, 0.01 / B
        bst = []
        out = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts.to(dtype), (eps.to(dtype), pp), gg, bst)
        return dict(y=res[0][-1], x0_bar=out[0][0], eps_bar=out[2], p_bar=spec.flatten_params(out[3]), ts_bar=out[1],
                    fwd=fst["n_iters"], bwd=bst[0]["n_iters"])

    def gpu(tol):
        xg = x.to(dev).requires_grad_(True)
        eg = eps.to(dev).requires_grad_(True)
        pf = spec.flatten_params({k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p.items()}).clone().requires_grad_(True)
        zc = z.to(dev)
        (ys, dl, rs), nfe = eno.odeint_sepaux(spec.fwd, spec.aug_dynamics, (xg, zc, zc), ts, eg, pf, rtol=tol, atol=tol)
        loss = (ys[-1] * gy.to(dev)).sum() + dl[-1].sum() / B + 0.01 * rs[-1].sum() / B
        loss.backward()
        return dict(y=ys[-1], x0_bar=xg.grad, eps_bar=eg.grad, p_bar=pf.grad, fwd=eno.last_stats["fwd"]["n_iters"],
                    bwd=eno.last_stats["bwd"][0]["n_iters"])

    truth = oracle(torch.float64, 1e-10)
    print(f"B {B} D {D} H {H}; truth (fp64, 1e-10): fwd {truth['fwd']} bwd {truth['bwd']} iterations")
    for tol in (1e-3, 1e-5, 1e-6, 1.4e-8):
        o, gq = oracle(torch.float32, tol), gpu(tol)
        print(f"tol {tol:g}: iterations fwd gpu/oracle {gq['fwd']}/{o['fwd']}  bwd {gq['bwd']}/{o['bwd']}")
        for k in ("y", "x0_bar", "eps_bar", "p_bar"):
            print(f"   {k:8s} gpu-oracle {rel(gq[k], o[k]):.2e}   gpu-truth {rel(gq[k], truth[k]):.2e}   oracle-truth {rel(o[k], truth[k]):.2e}")


if __name__ == "__main__":
    main()
