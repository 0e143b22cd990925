"""CPU: the oracle's restatement of the fixed-grid RK4 family (oracle/ode_oracle.py: rk4_solve, odeint_grid*,
rk4_odeint_rev) against tests/golden/css_grid.npz, the output of the unmodified reference solver (lib/ode.py:864-946,
1531-1593; tests/golden/make_grid.py)."""
import numpy as np
import pytest
import torch

from oracle import ffjord_oracle as fo, ode_oracle as oo
from tests.util import load


def _flat(tree):
    return oo.ravel(tree)[0].numpy()


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("reg_type", ["our", "fin"])
def test_oracle_reproduces_reference_grid_solver(tag, reg_type):
    g = load("css_grid.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    p = fo.init_css_params(D, H, seed=int(g["seed"]), scale=float(g["scale"]))
    x0, eps, gy = (torch.as_tensor(g[k]) for k in ("x0", "eps", "gy"))
    step = float(g[f"{tag}_step"])
    ts = torch.tensor([0., 1.])
    n_aux = 2 if reg_type == "our" else 3
    key = tag if reg_type == "our" else f"{tag}_fin"
    cl = fo.make_ffjord_closures("r2" if reg_type == "our" else "none", reg_type)
    z = torch.zeros(B, 1)
    res, nfe = oo.odeint_grid_split_fwd(cl["fwd"], (x0,) + (z,) * n_aux, ts, eps, p, n_aux=n_aux, step_size=step)
    assert nfe == float(g[f"{key}_nfe"]) == 16.0          # 4 steps of RK4: 0.25 x 4, or 0.3 x 3 + a clipped 0.1
    np.testing.assert_allclose(res[0].numpy(), g[f"{key}_y"], rtol=0, atol=1e-6)
    assert all(float(r.abs().max()) == 0.0 and tuple(r.shape) == (2, B, 1) for r in res[1:])
    gg = tuple(torch.zeros_like(r) for r in res)
    w = [1.0 / B, 0.01 / B, 0.02 / B]
    gg[0][-1] = gy
    for q in range(n_aux):
        gg[1 + q][-1] = w[q]
    out = oo.odeint_grid_split_rev(cl["aug_dynamics"], step, res, ts, (eps, p), gg)
    np.testing.assert_allclose(out[0][0].numpy(), g[f"{key}_x0_bar"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out[1].numpy(), g[f"{key}_ts_bar"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out[2].numpy(), g[f"{key}_eps_bar"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(_flat(out[3]), g[f"{key}_params_bar"], rtol=0, atol=2e-6)


def test_plain_grid_solve_and_step_clipping():
    g = load("css_grid.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    p = fo.init_css_params(D, H, seed=int(g["seed"]), scale=float(g["scale"]))
    cl = fo.make_ffjord_closures("none", "fin")
    x0, eps = torch.as_tensor(g["x0"]), torch.as_tensor(g["eps"])
    for tag in ("a", "b"):
        ys, nfe = oo.odeint_grid(cl["fwd"], x0, torch.tensor([0., 1.]), eps, p, step_size=float(g[f"{tag}_step"]))
        assert nfe == float(g[f"{tag}_plain_nfe"])
        np.testing.assert_allclose(ys.numpy(), g[f"{tag}_plain_y"], rtol=0, atol=1e-6)
    # a step larger than the interval is one clipped step; a tiny interval remainder still costs a whole step
    _, nfe = oo.odeint_grid(cl["fwd"], x0, torch.tensor([0., 1.]), eps, p, step_size=5.0)
    assert nfe == 4.0
    _, nfe = oo.odeint_grid(cl["fwd"], x0, torch.tensor([0., 1.]), eps, p, step_size=0.499)
    assert nfe == 12.0
