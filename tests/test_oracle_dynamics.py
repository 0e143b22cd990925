"""Cross-checks for the parts of the oracle that no reference vector can pin (jax.experimental.jet
is third-party and absent): Taylor-mode via nested jvp vs the closed form vs finite differences of
the true flow; solver known-answer tests from lib/ode.py:1782-1851; adjoint vs autograd-through-RK4."""
import math

import numpy as np
import pytest
import torch

from oracle import dynamics_oracle as do, ode_oracle as oo


def _case(dtype=torch.float64, B=3, D=8, H=4, scale=2.0):
    p = do.init_mlp_params(D, H, seed=3, dtype=dtype, scale=scale)
    g = torch.Generator().manual_seed(7)
    return p, torch.rand((B, D), generator=g, dtype=dtype)


@pytest.mark.parametrize("reg,K", [("r2", 2), ("r3", 3)])
def test_nested_jvp_equals_closed_form(reg, K):
    p, z = _case()
    t = torch.tensor(0.3, dtype=torch.float64)
    f = lambda zz, tt: do.mlp_dynamics(p, zz, tt)
    dz, dk = do.sol_recursive(f, z, t, reg)
    cf = do.mlp_taylor_closed(p, z, 0.3, K)
    assert (dz - cf[0]).abs().max() < 1e-13
    assert (dk - cf[K - 1]).abs().max() < 1e-12


def test_third_derivative_matches_finite_difference_of_flow():
    """d^3 z/dt^3 along the TRUE trajectory (RK4 with tiny steps), central differences of z'."""
    p, z = _case(scale=1.5)
    f = lambda zz, tt: do.mlp_dynamics(p, zz, tt)

    def flow(z0, t0, t1, n=200):
        h = (t1 - t0) / n
        zz, tt = z0, t0
        for _ in range(n):
            k1 = f(zz, tt); k2 = f(zz + h / 2 * k1, tt + h / 2); k3 = f(zz + h / 2 * k2, tt + h / 2); k4 = f(zz + h * k3, tt + h)
            zz = zz + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4); tt = tt + h
        return zz
    t0, e = 0.3, 1e-2
    zp, zm = flow(z, t0, t0 + e), flow(z, t0, t0 - e)
    d1 = lambda zz, tt: f(zz, torch.tensor(tt, dtype=torch.float64))
    # z''' = d^2/dt^2 (z') ~ (z'(t+e) - 2 z'(t) + z'(t-e)) / e^2
    fd3 = (d1(zp, t0 + e) - 2 * d1(z, t0) + d1(zm, t0 - e)) / e ** 2
    _, d3 = do.sol_recursive(f, z, torch.tensor(t0, dtype=torch.float64), "r3")
    assert (fd3 - d3).abs().max() / d3.abs().max() < 1e-3
    fd2 = (d1(zp, t0 + e) - d1(zm, t0 - e)) / (2 * e)
    _, d2 = do.sol_recursive(f, z, torch.tensor(t0, dtype=torch.float64), "r2")
    assert (fd2 - d2).abs().max() / d2.abs().max() < 1e-3


def test_linear_kat_against_expm():
    """lib/ode.py:1812-1835: y' = A y, A = U - U^T, y(t) = expm(A t) y0."""
    import scipy.linalg
    rng = np.random.default_rng(0)
    U = 0.1 * rng.standard_normal((10, 10))
    A = torch.as_tensor(U - U.T)
    y0 = torch.ones(10, dtype=torch.float64)
    ts = torch.tensor([0., 2., 5.], dtype=torch.float64)
    ys, nfe, st = oo.solve(lambda y, t: A @ y, y0, ts, 1e-9, 1e-9)
    for i, t in enumerate(ts):
        want = scipy.linalg.expm(A.numpy() * float(t)) @ y0.numpy()
        np.testing.assert_allclose(ys[i].numpy(), want, rtol=1e-6, atol=1e-7)
    assert nfe == 2 + 6 * st["n_iters"]


def test_against_scipy_rk45_on_pendulum():
    import scipy.integrate

    def pend(t, y):
        return [y[1], -0.25 * y[1] - 9.8 * math.sin(y[0])]
    sol = scipy.integrate.solve_ivp(pend, (0, 3), [math.pi - 0.1, 0.0], rtol=1e-10, atol=1e-10)
    f = lambda y, t: torch.stack([y[1], -0.25 * y[1] - 9.8 * torch.sin(y[0])])
    ys, _, _ = oo.solve(f, torch.tensor([math.pi - 0.1, 0.0], dtype=torch.float64),
                        torch.tensor([0., 3.], dtype=torch.float64), 1e-9, 1e-9)
    np.testing.assert_allclose(ys[-1].numpy(), sol.y[:, -1], rtol=1e-6, atol=1e-7)


def test_adjoint_gradients_against_autograd_through_fixed_rk4():
    """fp64: adjoint of the adaptive solve vs back-propagation through an unrolled fine RK4."""
    p, z = _case(B=2, D=8, H=4, scale=1.0)
    eps = torch.ones_like(z)
    cl = do.make_mnist_closures("r3")
    ts = torch.tensor([0., 1.], dtype=torch.float64)
    tol = 1e-10
    res, _ = oo.odeint_split_fwd(cl["fwd"], (z, torch.zeros(2, dtype=torch.float64)), ts, eps, p, rtol=tol, atol=tol)
    gy = torch.linspace(-1, 1, z.numel(), dtype=torch.float64).reshape(z.shape)
    gg = (torch.zeros_like(res[0]), torch.zeros_like(res[1]))
    gg[0][-1] = gy
    gg[1][-1] = 0.3
    out = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, p), gg)

    leaves = [p[do.L1]["b"], p[do.L1]["w"], p[do.L2]["b"], p[do.L2]["w"]]
    leaves = [l.clone().requires_grad_(True) for l in leaves]
    pp = {do.L1: {"b": leaves[0], "w": leaves[1]}, do.L2: {"b": leaves[2], "w": leaves[3]}}
    z0 = z.clone().requires_grad_(True)
    n = 200
    h = 1.0 / n
    y, r = z0, torch.zeros(2, dtype=torch.float64)

    def F(yy, tt):
        f, y1, y2 = do.mlp_taylor_closed(pp, yy, tt, 3)
        return f, (y2 ** 2).mean(dim=1)
    for i in range(n):
        t = i * h
        k1 = F(y, t); k2 = F(y + h / 2 * k1[0], t + h / 2); k3 = F(y + h / 2 * k2[0], t + h / 2); k4 = F(y + h * k3[0], t + h)
        y = y + h / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
        r = r + h / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
    loss = (y * gy).sum() + 0.3 * r.sum()
    grads = torch.autograd.grad(loss, [z0] + leaves)
    assert (out[0][0] - grads[0]).abs().max() / grads[0].abs().max() < 1e-6
    got = [out[3][do.L1]["b"], out[3][do.L1]["w"], out[3][do.L2]["b"], out[3][do.L2]["w"]]
    for a, b in zip(got, grads[1:]):
        assert (a - b).abs().max() / b.abs().max() < 1e-6
