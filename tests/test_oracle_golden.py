"""The oracle (oracle/ode_oracle.py) against the golden vectors that the UNMODIFIED reference solver
produced (tests/golden/make_golden.py).  Same backend (torch CPU), so agreement is expected to the
last bit for the solver core; a small tolerance is kept for library-version drift."""
import math

import numpy as np
import pytest
import torch

from oracle import dynamics_oracle as do, ode_oracle as oo
from tests.util import load, params_from_flat, flat_from_params

TIGHT = dict(rtol=2e-6, atol=1e-7)
TABLEAUX = ["heun", "fehlberg", "bosh", "rk_fehlberg", "cash_karp", "owrenzen", "owrenzen5", "tanyam"]


def test_kat_const_problem():
    g = load("kat_const.npz")
    a, b = 0.2, 3.0
    f = lambda y, t: a + (y - (a * t + b)) ** 5
    ts = torch.tensor([1., 8.])
    y0 = torch.tensor([a + b])
    ys, nfe, _ = oo.solve(f, y0, ts, 1e-5, 1e-5)
    assert nfe == float(g["dopri5_nfe"])
    np.testing.assert_allclose(ys.numpy(), g["dopri5_y"], **TIGHT)
    assert abs(float(ys[-1]) - (a * 8 + b)) < 1e-5            # exact solution y = a t + b
    for name in TABLEAUX:
        ys, nfe, _ = oo.solve(f, y0, ts, 1e-5, 1e-5, solver=name)
        assert nfe == float(g[f"{name}_nfe"]), name
        np.testing.assert_allclose(ys.numpy(), g[f"{name}_y"], **TIGHT)


@pytest.fixture(scope="module")
def fwd_case():
    g = load("mlp_forward.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    return g, params_from_flat(g["params"], D, H), torch.as_tensor(g["x0"])


@pytest.mark.parametrize("tag", ["d", "5", "3"])
def test_dopri5_forward(fwd_case, tag):
    g, params, x0 = fwd_case
    tol = float(g[f"fwd_{tag}_tol"])
    ys, nfe = oo.odeint(lambda y, t, p: do.mlp_dynamics(p, y, t), x0, torch.tensor([0., 1.]), params, rtol=tol, atol=tol)
    assert nfe == float(g[f"fwd_{tag}_nfe"])
    np.testing.assert_allclose(ys.numpy(), g[f"fwd_{tag}_y"], **TIGHT)


def test_dopri5_multi_output_and_mxstep(fwd_case):
    g, params, x0 = fwd_case
    f = lambda y, t, p: do.mlp_dynamics(p, y, t)
    ys, nfe = oo.odeint(f, x0, torch.as_tensor(g["fwd_multi_ts"]), params, rtol=1e-6, atol=1e-6)
    assert nfe == float(g["fwd_multi_nfe"])
    np.testing.assert_allclose(ys.numpy(), g["fwd_multi_y"], **TIGHT)
    ys, nfe = oo.odeint(f, x0, torch.tensor([0., 0.07]), params, rtol=1e-7, atol=1e-7, mxstep=2)
    assert nfe == float(g["fwd_mx2_nfe"])
    np.testing.assert_allclose(ys.numpy(), g["fwd_mx2_y"], **TIGHT)


@pytest.mark.parametrize("name", TABLEAUX)
def test_tableau_sweep(fwd_case, name):
    g, params, x0 = fwd_case
    fun = lambda y, t: do.mlp_dynamics(params, y, t).reshape(-1)
    for tag, tol in [("3", 1e-3), ("5", 1e-5)]:
        if f"{name}_{tag}_y" not in g:
            continue
        ys, nfe, _ = oo.solve(fun, x0.reshape(-1), torch.tensor([0., 1.]), tol, tol, solver=name)
        assert nfe == float(g[f"{name}_{tag}_nfe"]), (name, tag)
        np.testing.assert_allclose(ys.numpy(), g[f"{name}_{tag}_y"], **TIGHT)


@pytest.mark.parametrize("reg", ["r2", "r3"])
def test_split_solve_and_adjoint(reg):
    g = load("mlp_adjoint.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    params = params_from_flat(g["params"], D, H)
    x0, eps, gy = (torch.as_tensor(g[k]) for k in ("x0", "eps", "gy"))
    cl = do.make_mnist_closures(reg)
    ts = torch.tensor([0., 1.])
    tol = 1e-5
    res, nfe = oo.odeint_split_fwd(cl["fwd"], (x0, torch.zeros(B)), ts, eps, params, n_aux=1, rtol=tol, atol=tol)
    k = f"{reg}_5"
    assert nfe == float(g[f"{k}_nfe"])
    np.testing.assert_allclose(res[0][-1].numpy(), g[f"{k}_y1"], **TIGHT)
    gg = (torch.zeros_like(res[0]), torch.zeros_like(res[1]))
    gg[0][-1] = gy
    gg[1][-1] = float(g["g_r"])
    out = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, params), gg)
    np.testing.assert_allclose(out[0][0].numpy(), g[f"{k}_x0_bar"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[1].numpy(), g[f"{k}_ts_bar"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[2].numpy(), g[f"{k}_eps_bar"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(flat_from_params(out[3]).numpy(), g[f"{k}_params_bar"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["5", "m"])
def test_sepaux_fin_solve_and_adjoint(tag):
    """odeint_sepaux with the 'fin' aug_dynamics (lib/ode.py:678-697, 1686-1693; mnist.py:294-313):
    eps is read by the RHS, so eps_bar is a live block of the adjoint state."""
    g = load("mlp_adjoint_fin.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    params = params_from_flat(g["params"], D, H)
    x0, eps, gy = (torch.as_tensor(g[k]) for k in ("x0", "eps", "gy"))
    cl = do.make_mnist_closures("none", reg_type="fin")
    k = f"fin_{tag}"
    ts, tol = torch.as_tensor(g[f"{k}_ts"]), float(g[f"{k}_tol"])
    res, nfe = oo.odeint_split_fwd(cl["fwd"], (x0, torch.zeros(B), torch.zeros(B)), ts, eps, params, n_aux=2,
                                   rtol=tol, atol=tol)
    assert nfe == float(g[f"{k}_nfe"])
    np.testing.assert_allclose(res[0].numpy(), g[f"{k}_y"], **TIGHT)
    gg = tuple(torch.zeros_like(r) for r in res)
    if tag == "m":
        for i in range(len(ts)):
            gg[0][i] = gy * (0.5 + i)
            gg[1][i] = 0.01 * (1 + i)
            gg[2][i] = 0.02 * (3 - i)
    else:
        gg[0][-1], gg[1][-1], gg[2][-1] = gy, float(g["g_fro"]), float(g["g_kin"])
    out = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, params), gg)
    np.testing.assert_allclose(out[0][0].numpy(), g[f"{k}_x0_bar"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[1].numpy(), g[f"{k}_ts_bar"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[2].numpy(), g[f"{k}_eps_bar"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(flat_from_params(out[3]).numpy(), g[f"{k}_params_bar"], rtol=1e-5, atol=1e-6)


def test_per_example_solves_golden():
    """mnist.py:350-353 (--vmap): every sample as its own problem; includes an mxstep-limited (extrapolated) run."""
    g = load("mlp_vmap.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    params = params_from_flat(g["params"], D, H)
    x0 = torch.as_tensor(g["x0"])
    ts = torch.tensor([0., 1.])
    for tag, mx in [("5", math.inf), ("mx2", 2)]:
        tol = float(g[f"vmap_{tag}_tol"])
        for b in range(B):
            ys, nfe = oo.odeint(do.mlp_dynamics_xtp, x0[b:b + 1], ts, params, rtol=tol, atol=tol, mxstep=mx)
            assert nfe == float(g[f"vmap_{tag}_nfe"][b])
            np.testing.assert_allclose(ys[:, 0].numpy(), g[f"vmap_{tag}_y"][:, b], **TIGHT)


def test_pendulum_adjoint_multi_segment():
    g = load("pendulum.npz")

    def pend(y, t, m, gr):
        return torch.stack([y[1], -m * y[1] - gr * torch.sin(y[0])])
    y0 = torch.tensor([math.pi - 0.1, 0.0])
    ts = torch.as_tensor(g["ts"])
    args = (torch.tensor(0.25), torch.tensor(9.8))
    ys, nfe, _ = oo.solve(lambda y, t: pend(y, t, *args), y0, ts, 1e-6, 1e-6)
    assert nfe == float(g["nfe"])
    np.testing.assert_allclose(ys.numpy(), g["ys"], **TIGHT)
    out = oo.odeint_rev(pend, 1e-6, 1e-6, math.inf, ys, ts, args, torch.ones_like(ys))
    np.testing.assert_allclose(out[0].numpy(), g["y0_bar"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[1].numpy(), g["ts_bar"], rtol=1e-5, atol=1e-6)
    assert abs(float(out[2]) - float(g["m_bar"])) < 1e-4 and abs(float(out[3]) - float(g["g_bar"])) < 1e-4


def test_stiff_case_has_rejected_steps_and_matches_reference():
    """tests/golden/mlp_stiff.npz: solves whose dopri5 loops REJECT steps, forward and adjoint (lib/ode.py:843-848).
    The oracle reproduces the unmodified reference on them and its statistics prove the rejections are there."""
    g = load("mlp_stiff.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    params = params_from_flat(g["params"], D, H)
    x0, eps, gy = (torch.as_tensor(g[k]) for k in ("x0", "eps", "gy"))
    ys, nfe, st = oo.solve(lambda y, t: do.mlp_dynamics(params, y.reshape(B, D), t).reshape(-1), x0.reshape(-1),
                           torch.as_tensor(g["fwd3_ts"]), 1e-6, 1e-6)
    assert nfe == float(g["fwd3_nfe"]) and st["n_accepted"] < st["n_iters"]
    np.testing.assert_allclose(ys.reshape(-1, B, D).numpy(), g["fwd3_y"], **TIGHT)
    ts = torch.tensor([0., 1.])
    cl = do.make_mnist_closures("r3")
    for tag in ("5", "6"):
        k, tol = f"r3_{tag}", float(g[f"r3_{tag}_tol"])
        res, nfe, fst = oo.odeint_split_fwd(cl["fwd"], (x0, torch.zeros(B)), ts, eps, params, n_aux=1, rtol=tol, atol=tol,
                                            return_stats=True)
        assert nfe == float(g[f"{k}_nfe"]) and fst["n_accepted"] < fst["n_iters"]
        np.testing.assert_allclose(res[0][-1].numpy(), g[f"{k}_y1"], **TIGHT)
        gg = (torch.zeros_like(res[0]), torch.zeros_like(res[1]))
        gg[0][-1], gg[1][-1] = gy, float(g["g_r"])
        bst = []
        out = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, params), gg, bst)
        assert bst[0]["n_accepted"] < bst[0]["n_iters"]
        np.testing.assert_allclose(out[0][0].numpy(), g[f"{k}_x0_bar"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out[1].numpy(), g[f"{k}_ts_bar"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(flat_from_params(out[3]).numpy(), g[f"{k}_params_bar"], rtol=1e-5, atol=1e-6)
    clf = do.make_mnist_closures("none", reg_type="fin")
    tol = float(g["fin_tol"])
    res, nfe, fst = oo.odeint_split_fwd(clf["fwd"], (x0, torch.zeros(B), torch.zeros(B)), ts, eps, params, n_aux=2, rtol=tol,
                                        atol=tol, return_stats=True)
    assert nfe == float(g["fin_nfe"]) and fst["n_accepted"] < fst["n_iters"]
    gg = tuple(torch.zeros_like(r) for r in res)
    gg[0][-1], gg[1][-1], gg[2][-1] = gy, float(g["g_fro"]), float(g["g_kin"])
    bst = []
    out = oo.odeint_split_rev(clf["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, params), gg, bst)
    assert bst[0]["n_accepted"] < bst[0]["n_iters"]
    np.testing.assert_allclose(out[2].numpy(), g["fin_eps_bar"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(flat_from_params(out[3]).numpy(), g["fin_params_bar"], rtol=1e-5, atol=1e-6)
