"""Generates tests/golden/conv_ffjord.npz: the UNMODIFIED reference solver (/root/reference/lib/ode.py on
oracle/jax_shim.py) on the forward closures of ffjord_mnist.py with ConcatConv2D dynamics (oracle/ffjord_mnist_oracle.py)
at a small size: the nfe_fn solve (ffjord_dynamics on (y, delta_logp), ffjord_mnist.py:362-385) and the forward_all solve
(all_aug_dynamics on (y, delta_logp, r, fro, kin), 326-338, 409-417).  Run in the build container only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.jax_shim import load_reference_ode, ravel_pytree  # noqa: E402
from oracle import ffjord_mnist_oracle as fm  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = load_reference_ode()
    B, img, hidden, seed, scale, tol = 3, (6, 6, 1), 32, 2, 1.5, 1e-5
    params = fm.init_conv_params(hidden=hidden, seed=seed, scale=scale, last_scale=scale)
    g = torch.Generator().manual_seed(12)
    x0 = torch.randn((B,) + img, generator=g)
    eps = torch.randint(0, 2, (B,) + img, generator=g).float() * 2 - 1          # Rademacher, ffjord_mnist.py:147-152
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    out = {}
    for reg in ("none", "r2"):
        cl = fm.make_ffjord_mnist_closures(reg, "our", img)
        (ys, dl), nfe = ref.odeint(cl["ffjord_dynamics"], (x0, z), ts, eps, params, rtol=tol, atol=tol)
        out[f"ffjord_{reg}_y"], out[f"ffjord_{reg}_dlogp"], out[f"ffjord_{reg}_nfe"] = ys.numpy(), dl.numpy(), float(nfe)
    cl = fm.make_ffjord_mnist_closures("r2", "our", img)
    (ys, dl, rs, fro, kin), nfe = ref.odeint(cl["all_aug_dynamics"], (x0, z, z, z, z), ts, eps, params, rtol=tol, atol=tol)
    out.update(all_y=ys.numpy(), all_dlogp=dl.numpy(), all_r=rs.numpy(), all_fro=fro.numpy(), all_kin=kin.numpy(), all_nfe=float(nfe))
    ys, nfe = ref.odeint(cl["dynamics_wrap"], x0, ts, params, rtol=tol, atol=tol)
    out.update(plain_y=ys.numpy(), plain_nfe=float(nfe))
    np.savez(os.path.join(OUT, "conv_ffjord.npz"), B=B, img=np.array(img), hidden=hidden, seed=seed, scale=scale, tol=tol,
             x0=x0.numpy(), eps=eps.numpy(), params=ravel_pytree(params)[0].numpy(), **out)
    print({k: (v if np.isscalar(v) else getattr(v, "shape", v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
