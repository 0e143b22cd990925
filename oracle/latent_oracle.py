"""CPU oracle for the generative ODE of latent_ode.py (BASELINE config C3) - groundwork for the next rows of the
scope table, not yet backed by kernels.

TEST INFRASTRUCTURE.  Restates, in eager torch:
  GenDynamics ............. latent_ode.py:191-208   repeat (tanh, Linear(units)) x (layers+1), then tanh, Linear(latent)
                                                    (no time input; the state is reshaped to [-1, latent_dim])
  augment_dynamics ........ latent_ode.py:238-256   (dz/dt, mean(r^2)) with r = d^K z/dt^K, one scalar per call
  aug_init ................ latent_ode.py:337-355   (z, 0.)
The solver call it feeds is ``jax.vmap(lambda z, t: odeint(dynamics, aug_init(z), t, params))`` over trajectories
(latent_ode.py:344-350): every trajectory is its own differentiable odeint problem with ~49 output times, i.e.
the per-example structure of csrc/cl_vmap.cuh plus the multi-segment adjoint.

PARITY UNPINNED for the Taylor-mode part (jax.experimental.jet, see dynamics_oracle.py).
Nothing under easy-neural-ode_b200/ imports this module.
"""
import math

import torch

from .dynamics_oracle import sol_recursive


def layer_names(n_linear):
    return ["gen_dynamics/linear" + ("" if i == 0 else f"_{i}") for i in range(n_linear)]


def init_gen_params(latent_dim=20, units=50, layers=3, seed=0, dtype=torch.float32, scale=1.0):
    """layers + 1 hidden Linear(units) and one Linear(latent_dim); Haiku-default init (GenDynamics is built
    without init_kwargs, latent_ode.py:293-296)."""
    g = torch.Generator().manual_seed(seed)

    def tn(shape, fan_in):
        w = torch.empty(shape, dtype=torch.float64)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
        return (w * (scale / math.sqrt(fan_in))).to(dtype)

    dims = [latent_dim] + [units] * (layers + 1) + [latent_dim]
    return {n: {"w": tn((a, b), a), "b": torch.zeros(b, dtype=dtype)}
            for n, (a, b) in zip(layer_names(len(dims) - 1), zip(dims[:-1], dims[1:]))}


def gen_dynamics(params, y, t=None):
    """f(y) of latent_ode.py:202-208 (time-independent)."""
    names = layer_names(len(params))
    x = y.reshape(-1, params[names[0]]["w"].shape[0])
    for n in names:
        x = torch.tanh(x) @ params[n]["w"] + params[n]["b"]
    return x


def gen_taylor_closed(params, y, order):
    """Closed-form z^(1..order) (order <= 3) along z' = f(z), SURVEY.md Appendix B:
    a = tanh(x): a' = p1 x', a'' = p2 x'^2 + p1 x'' with p1 = 1 - a^2, p2 = -2 a p1; linear maps act
    coefficient-wise, biases only on the primal."""
    names = layer_names(len(params))
    x0 = y.reshape(-1, params[names[0]]["w"].shape[0])
    f = gen_dynamics(params, y)
    outs = [f]
    x1 = f
    x2 = None
    if order >= 2:
        z, z1 = x0, x1
        for n in names:
            a = torch.tanh(z)
            z, z1 = a @ params[n]["w"] + params[n]["b"], ((1 - a * a) * z1) @ params[n]["w"]
        outs.append(z1)
        x2 = z1
    if order >= 3:
        z, z1, z2 = x0, x1, x2
        for n in names:
            a = torch.tanh(z)
            p1 = 1 - a * a
            p2 = -2 * a * p1
            w = params[n]["w"]
            z, z1, z2 = a @ w + params[n]["b"], (p1 * z1) @ w, (p2 * z1 * z1 + p1 * z2) @ w
        outs.append(z2)
    return outs


def make_latent_closures(reg="r3"):
    def gen_dynamics_wrap(x, t, params):
        return gen_dynamics(params, x, t)

    def aug_dynamics(yr, t, params):
        y = yr[0]
        y0, r = sol_recursive(lambda _y, _t: gen_dynamics_wrap(_y, _t, params), y, t, reg)
        return y0, torch.mean(r ** 2)

    return dict(gen_dynamics_wrap=gen_dynamics_wrap, aug_dynamics=aug_dynamics)
