"""Oracle groundwork for BASELINE config C2 (ffjord_tabular.py): the restated closures are cross-checked
independently (closed-form Taylor coefficients, exact Jacobian), and the split solve odeint_sepaux with them is
pinned to golden vectors produced by the unmodified reference solver (tests/golden/make_golden.py)."""
import math

import numpy as np
import pytest
import torch

from oracle import ffjord_oracle as fo, ode_oracle as oo
from oracle.dynamics_oracle import sol_recursive
from tests.util import load

torch.manual_seed(0)


def _case(dtype=torch.float64, B=5, D=6, H=12):
    p = fo.init_css_params(D, H, seed=3, dtype=dtype, scale=1.5)
    g = torch.Generator().manual_seed(9)
    x = torch.randn((B, D), generator=g, dtype=dtype)
    eps = torch.randn((B, D), generator=g, dtype=dtype)
    return p, x, eps


def test_softplus_is_logaddexp_with_zero():
    x = torch.linspace(-30, 30, 61, dtype=torch.float64)
    assert torch.allclose(fo.softplus(x), torch.logaddexp(x, torch.zeros_like(x)), rtol=1e-14, atol=0)


def test_second_time_derivative_closed_form_equals_nested_jvp():
    p, x, _ = _case()
    f, y1 = fo.css_taylor_closed(p, x, 0.3)
    f_w, y1_w = sol_recursive(lambda y, t: fo.css_dynamics(p, y, t), x, 0.3, "r2")
    assert torch.allclose(f, f_w, rtol=1e-13, atol=1e-14)
    assert torch.allclose(y1, y1_w, rtol=1e-12, atol=1e-13)


def test_hutchinson_term_equals_eps_J_eps_from_the_full_jacobian():
    p, x, eps = _case()
    cl = fo.make_ffjord_closures()
    dy, neg_div = cl["ffjord_dynamics"]((x, torch.zeros(x.shape[0], 1, dtype=x.dtype)), 0.7, eps, p)
    for b in range(x.shape[0]):
        J = torch.autograd.functional.jacobian(lambda v: fo.css_dynamics(p, v[None], 0.7)[0], x[b])
        assert abs(float(-neg_div[b, 0]) - float(eps[b] @ J @ eps[b])) < 1e-12
    assert torch.allclose(dy, fo.css_dynamics(p, x, 0.7))


def test_aug_dynamics_shapes_follow_the_reference():
    p, x, eps = _case(torch.float32)
    z = torch.zeros(x.shape[0], 1)
    our = fo.make_ffjord_closures("r2", "our")["aug_dynamics"]((x, z, z), 0.2, eps, p)
    fin = fo.make_ffjord_closures("r2", "fin")["aug_dynamics"]((x, z, z, z), 0.2, eps, p)
    assert [tuple(o.shape) for o in our] == [(5, 6), (5, 1), (5,)]
    assert [tuple(o.shape) for o in fin] == [(5, 6), (5, 1), (5,), (5,)]


def test_sepaux_solve_and_adjoint_golden():
    g = load("css_sepaux.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    p = fo.init_css_params(D, H, seed=int(g["seed"]), scale=float(g["scale"]))
    x0, eps, gy = (torch.as_tensor(g[k]) for k in ("x0", "eps", "gy"))
    cl = fo.make_ffjord_closures("r2", "our")
    ts = torch.tensor([0., 1.])
    tol = float(g["tol"])
    y0 = (x0, torch.zeros(B, 1), torch.zeros(B, 1))
    res, nfe = oo.odeint_split_fwd(cl["fwd"], y0, ts, eps, p, n_aux=2, rtol=tol, atol=tol)
    assert nfe == float(g["nfe"])
    np.testing.assert_allclose(res[0].numpy(), g["y"], rtol=1e-6, atol=1e-7)
    gg = tuple(torch.zeros_like(r) for r in res)
    gg[0][-1], gg[1][-1], gg[2][-1] = gy, 1.0 / B, 0.01 / B
    out = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, p), gg)
    np.testing.assert_allclose(out[0][0].numpy(), g["x0_bar"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[2].numpy(), g["eps_bar"], rtol=1e-5, atol=1e-6)
    flat = torch.cat([v.reshape(-1) for n in fo.layer_names(len(p)) for _, v in sorted(out[3][n].items())])
    np.testing.assert_allclose(flat.numpy(), g["params_bar"], rtol=1e-5, atol=1e-6)
