"""Generates tests/golden/latent_traj.npz by running the UNMODIFIED reference solver (/root/reference/lib/ode.py:
odeint -> _dopri5_odeint forward, _odeint_rev backward) on the torch-backed JAX stand-in (oracle/jax_shim.py), in
float64 like latent_ode.py (jax_enable_x64, latent_ode.py:23), for a few trajectories of the generative ODE
(latent_ode.py:337-350: every trajectory is its own odeint problem with several output times; the cotangents arrive
at every output time for z and at the last one for r, latent_ode.py:362-366).

The dynamics closures come from oracle/latent_oracle.py (GenDynamics + augment_dynamics restated; jet is not importable).
Run in the build container only: ``python tests/golden/make_latent.py``.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.jax_shim import load_reference_ode, ravel_pytree  # noqa: E402
from oracle import latent_oracle as lo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_default_dtype(torch.float64)
    ref = load_reference_ode()
    D, U, layers, seed, scale, n_traj = 8, 12, 2, 4, 1.5, 3
    rec = dict(D=D, U=U, layers=layers, seed=seed, scale=scale, n_traj=n_traj)
    params = lo.init_gen_params(D, U, layers, seed=seed, scale=scale, dtype=torch.float64)
    g = torch.Generator().manual_seed(11)
    z0 = 0.5 * torch.randn((n_traj, D), generator=g, dtype=torch.float64)
    ts = torch.tensor([0., 0.2, 0.45, 0.8, 1.0], dtype=torch.float64)
    T = ts.numel()
    gz = torch.randn((n_traj, T, D), generator=g, dtype=torch.float64)
    rec.update(z0=z0.numpy(), ts=ts.numpy(), gz=gz.numpy(), params=ravel_pytree(params)[0].numpy())
    for reg, tol, g_r_last in [("r3", 1e-8, 0.01), ("r2", 1e-6, 0.05)]:
        cl = lo.make_latent_closures(reg)
        y0_tree = (z0[0:1], torch.zeros((), dtype=torch.float64))
        func = ref.ravel_first_arg(cl["aug_dynamics"], ravel_pytree(y0_tree)[1])
        zs_all, rs_all, nfe_all, z0b, r0b, tsb, pb = [], [], [], [], [], [], []
        for k in range(n_traj):
            y0 = (z0[k:k + 1], torch.zeros((), dtype=torch.float64))
            flat0 = ravel_pytree(y0)[0]
            (ys, nfe), res = ref._odeint_fwd(func, tol, tol, math.inf, flat0, ts, params)
            gflat = torch.zeros_like(ys)
            gflat[:, :D] = gz[k]
            gflat[-1, D] = g_r_last
            out = ref._odeint_rev(func, tol, tol, math.inf, res, (gflat, None))
            zs_all.append(ys[:, :D].numpy()); rs_all.append(ys[:, D].numpy()); nfe_all.append(float(nfe))
            z0b.append(out[0][:D].numpy()); r0b.append(float(out[0][D])); tsb.append(out[1].numpy())
            pb.append(ravel_pytree(out[2])[0].numpy())
        rec.update({f"{reg}_tol": tol, f"{reg}_g_r": g_r_last, f"{reg}_zs": np.stack(zs_all), f"{reg}_rs": np.stack(rs_all),
                    f"{reg}_nfe": np.array(nfe_all), f"{reg}_z0_bar": np.stack(z0b), f"{reg}_r0_bar": np.array(r0b),
                    f"{reg}_ts_bar": np.stack(tsb), f"{reg}_params_bar": np.stack(pb)})
        print(reg, "nfe", nfe_all)
    np.savez(os.path.join(OUT, "latent_traj.npz"), **rec)
    print("written", os.path.join(OUT, "latent_traj.npz"))


if __name__ == "__main__":
    main()
