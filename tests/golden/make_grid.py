"""Generates tests/golden/css_grid.npz by running the UNMODIFIED reference fixed-grid solver (/root/reference/lib/ode.py:
_rk4_odeint 864-903, its split wrappers 906-946 and the adjoint _rk4_odeint_rev 1531-1593) on the torch-backed JAX
stand-in (oracle/jax_shim.py), with the ConcatSquash FFJORD closures of oracle/ffjord_oracle.py.

Run in the build container only (``python tests/golden/make_grid.py``); the GPU box reads the committed .npz.
Two step sizes: 0.25 (four equal steps) and 0.3 (three steps + a clipped last one of 0.1).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.jax_shim import load_reference_ode, ravel_pytree  # noqa: E402
from oracle import ffjord_oracle as fo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = load_reference_ode()
    B, D, H, seed, scale = 5, 6, 12, 3, 1.5
    params = fo.init_css_params(D, H, seed=seed, scale=scale)
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn((B, D), generator=g)
    eps = torch.randn((B, D), generator=g)
    gy = torch.randn((B, D), generator=g)
    ts = torch.tensor([0., 1.])
    rec = dict(B=B, D=D, H=H, seed=seed, scale=scale, x0=x0.numpy(), eps=eps.numpy(), gy=gy.numpy())
    for tag, step in (("a", 0.25), ("b", 0.3)):
        # odeint_grid_sepaux_one: (y, delta_logp, r) with aug_dynamics 'our' (ffjord_tabular.py:77, 299-304)
        cl = fo.make_ffjord_closures("r2", "our")
        y0 = (x0, torch.zeros(B, 1), torch.zeros(B, 1))
        fwd_w = ref.ravel_first_arg(cl["fwd"], ravel_pytree(y0[0])[1])
        rev_w = ref.ravel_first_arg(cl["aug_dynamics"], ravel_pytree(y0)[1])
        (res, nfe), saved = ref._rk4_odeint_sepaux_one_fwd(fwd_w, rev_w, step, y0, ts, eps, params)
        gg = tuple(torch.zeros_like(r) for r in res)
        gg[0][-1], gg[1][-1], gg[2][-1] = gy, 1.0 / B, 0.01 / B
        out = ref._rk4_odeint_sepaux_one_rev(fwd_w, rev_w, step, saved, (gg, None))
        rec.update({f"{tag}_step": step, f"{tag}_y": res[0].numpy(), f"{tag}_nfe": float(nfe), f"{tag}_x0_bar": out[0][0].numpy(),
                    f"{tag}_ts_bar": out[1].numpy(), f"{tag}_eps_bar": out[2].numpy(),
                    f"{tag}_params_bar": ravel_pytree(out[3])[0].numpy()})
        # odeint_grid_sepaux: (y, delta_logp, fro, kin) with aug_dynamics 'fin' (ffjord_tabular.py:78)
        cl = fo.make_ffjord_closures("none", "fin")
        y0 = (x0, torch.zeros(B, 1), torch.zeros(B, 1), torch.zeros(B, 1))
        fwd_w = ref.ravel_first_arg(cl["fwd"], ravel_pytree(y0[0])[1])
        rev_w = ref.ravel_first_arg(cl["aug_dynamics"], ravel_pytree(y0)[1])
        (res, nfe), saved = ref._rk4_odeint_sepaux_fwd(fwd_w, rev_w, step, y0, ts, eps, params)
        gg = tuple(torch.zeros_like(r) for r in res)
        gg[0][-1], gg[1][-1], gg[2][-1], gg[3][-1] = gy, 1.0 / B, 0.01 / B, 0.02 / B
        out = ref._rk4_odeint_sepaux_rev(fwd_w, rev_w, step, saved, (gg, None))
        rec.update({f"{tag}_fin_y": res[0].numpy(), f"{tag}_fin_nfe": float(nfe), f"{tag}_fin_x0_bar": out[0][0].numpy(),
                    f"{tag}_fin_ts_bar": out[1].numpy(), f"{tag}_fin_eps_bar": out[2].numpy(),
                    f"{tag}_fin_params_bar": ravel_pytree(out[3])[0].numpy()})
        # plain odeint_grid
        ys, nfe = ref.odeint_grid(cl["fwd"], x0, ts, eps, params, step_size=step)
        rec.update({f"{tag}_plain_y": ys.numpy(), f"{tag}_plain_nfe": float(nfe)})
    np.savez(os.path.join(OUT, "css_grid.npz"), **rec)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in rec.items()})


if __name__ == "__main__":
    main()
