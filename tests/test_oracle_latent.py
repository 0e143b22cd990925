"""Oracle groundwork for BASELINE config C3 (latent_ode.py generative ODE): closed-form Taylor coefficients of the
tanh stack against nested forward-mode AD, and one trajectory's differentiable multi-output odeint (forward +
adjoint over 4 segments, latent_ode.py:344-350) against golden vectors from the unmodified reference solver."""
import math

import numpy as np
import torch

from oracle import latent_oracle as lo, ode_oracle as oo
from oracle.dynamics_oracle import sol_recursive
from tests.util import load


def test_taylor_closed_form_equals_nested_jvp():
    p = lo.init_gen_params(8, 12, 2, seed=1, dtype=torch.float64, scale=1.5)
    y = torch.randn((1, 8), generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    f, y1, y2 = lo.gen_taylor_closed(p, y, 3)
    fun = lambda z, t: lo.gen_dynamics(p, z)
    f_w, y1_w = sol_recursive(fun, y, 0.0, "r2")
    _, y2_w = sol_recursive(fun, y, 0.0, "r3")
    assert torch.allclose(f, f_w, rtol=1e-13, atol=1e-14)
    assert torch.allclose(y1, y1_w, rtol=1e-12, atol=1e-13)
    assert torch.allclose(y2, y2_w, rtol=1e-11, atol=1e-12)


def test_trajectory_solve_and_adjoint_golden():
    g = load("latent_gen.npz")
    D, U, Lh = int(g["D"]), int(g["U"]), int(g["layers"])
    p = lo.init_gen_params(D, U, Lh, seed=int(g["seed"]), scale=float(g["scale"]))
    z0 = torch.as_tensor(g["z0"])
    ts = torch.as_tensor(g["ts"])
    tol = float(g["tol"])
    cl = lo.make_latent_closures("r3")
    y0 = (z0, torch.zeros(()))
    (zs, rs), nfe = oo.odeint(cl["aug_dynamics"], y0, ts, p, rtol=tol, atol=tol)
    assert nfe == float(g["nfe"])
    np.testing.assert_allclose(zs.numpy(), g["zs"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rs.numpy(), g["rs"], rtol=1e-5, atol=1e-8)


def test_oracle_reproduces_reference_trajectories_fp64():
    """tests/golden/latent_traj.npz (unmodified reference odeint + _odeint_rev per trajectory, float64,
    tests/golden/make_latent.py) against the oracle restatement: forward outputs, NFE and the complete adjoint."""
    from oracle import ode_oracle as oo
    g = load("latent_traj.npz")
    D, U, Lh, n = int(g["D"]), int(g["U"]), int(g["layers"]), int(g["n_traj"])
    p = lo.init_gen_params(D, U, Lh, seed=int(g["seed"]), scale=float(g["scale"]), dtype=torch.float64)
    flat = torch.cat([torch.cat([p[k]["b"].reshape(-1), p[k]["w"].reshape(-1)]) for k in lo.layer_names(Lh + 2)])
    np.testing.assert_array_equal(flat.numpy(), g["params"])
    ts = torch.as_tensor(g["ts"])
    for reg in ("r3", "r2"):
        tol = float(g[f"{reg}_tol"])
        cl = lo.make_latent_closures(reg)
        for k in range(n):
            z0 = torch.as_tensor(g["z0"][k:k + 1])
            y0 = (z0, torch.zeros((), dtype=torch.float64))
            flat0, unravel = oo.ravel(y0)
            func = lambda yf, t, params: oo.ravel(cl["aug_dynamics"](unravel(yf), t, params))[0]
            ys, nfe = oo.odeint(func, flat0, ts, p, rtol=tol, atol=tol)
            assert nfe == float(g[f"{reg}_nfe"][k])
            np.testing.assert_allclose(ys[:, :D].numpy(), g[f"{reg}_zs"][k], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(ys[:, D].numpy(), g[f"{reg}_rs"][k], rtol=1e-12, atol=1e-16)
            gflat = torch.zeros_like(ys)
            gflat[:, :D] = torch.as_tensor(g["gz"][k])
            gflat[-1, D] = float(g[f"{reg}_g_r"])
            out = oo.odeint_rev(func, tol, tol, math.inf, ys, ts, (p,), gflat)
            np.testing.assert_allclose(out[0][:D].numpy(), g[f"{reg}_z0_bar"][k], rtol=1e-11, atol=1e-14)
            np.testing.assert_allclose(out[1].numpy(), g[f"{reg}_ts_bar"][k], rtol=1e-11, atol=1e-14)
            pb = torch.cat([torch.cat([out[2][nm]["b"].reshape(-1), out[2][nm]["w"].reshape(-1)]) for nm in lo.layer_names(Lh + 2)])
            np.testing.assert_allclose(pb.numpy(), g[f"{reg}_params_bar"][k], rtol=1e-10, atol=1e-14)
