"""CPU oracle for the closures of ffjord_tabular.py (BASELINE config C2) - groundwork for the next rows of the
scope table, not yet backed by kernels.

TEST INFRASTRUCTURE.  Restates, in eager torch:
  _logaddexp / softplus ... ffjord_tabular.py:92-103
  ConcatSquashLinear ...... ffjord_tabular.py:140-153   layer(x) * sigmoid(gate(t)) + hyper_bias(t)
  NN_Dynamics ............. ffjord_tabular.py:163-193   softplus between layers, none after the last
  sol_recursive ........... ffjord_tabular.py:116-136   fixed K = 2 whatever --reg says (two jet calls)
  reg_dynamics ............ ffjord_tabular.py:237-248
  ffjord(2)_dynamics ...... ffjord_tabular.py:250-268   jvp with eps: div = sum_d eps * (J eps), forward mode
  aug_dynamics ............ ffjord_tabular.py:270-284   (dy, -div, r) or (dy, -div, fro, kin)
  all_aug_dynamics ........ ffjord_tabular.py:286-297

PARITY UNPINNED for the Taylor-mode part (jax.experimental.jet, see dynamics_oracle.py): restated with nested
forward-mode AD and cross-checked by the closed form of SURVEY.md Appendix B (tests/test_oracle_ffjord.py).
Nothing under easy-neural-ode_b200/ imports this module.
"""
import math

import torch

from .dynamics_oracle import sigmoid, sol_recursive


def softplus(x):
    """_logaddexp(x, 0): max(x, 0) + log1p(exp(-|x|))."""
    return torch.clamp(x, min=0) + torch.log1p(torch.exp(-torch.abs(x)))


def layer_names(n_layers):
    """Haiku module names of NN_Dynamics' ConcatSquashLinear stack, in call order."""
    return ["nn__dynamics/concat_squash_linear" + ("" if i == 0 else f"_{i}") for i in range(n_layers)]


def init_css_params(dim, hidden, n_hidden_layers=2, seed=0, dtype=torch.float32, scale=1.0):
    """hidden_dims = (hidden,) * n_hidden_layers + (dim,).  Haiku-default Linear init (truncated normal, sigma =
    1/sqrt(fan_in), zero biases); the two hyper-networks take the scalar t, so their fan_in is 1."""
    g = torch.Generator().manual_seed(seed)

    def tn(shape, fan_in):
        w = torch.empty(shape, dtype=torch.float64)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
        return (w * (scale / math.sqrt(fan_in))).to(dtype)

    dims = [dim] + [hidden] * n_hidden_layers + [dim]
    params = {}
    for name, (d_in, d_out) in zip(layer_names(len(dims) - 1), zip(dims[:-1], dims[1:])):
        params[name] = {"w": tn((d_in, d_out), d_in), "b": torch.zeros(d_out, dtype=dtype),
                        "w_gate": tn((1, d_out), 1), "b_gate": torch.zeros(d_out, dtype=dtype),
                        "w_bias": tn((1, d_out), 1)}
    return params


def css_dynamics(params, x, t):
    """f(x, t) of ffjord_tabular.py:181-193; x: [B, D]."""
    names = layer_names(len(params))
    dx = x.reshape(-1, params[names[0]]["w"].shape[0])
    t = torch.as_tensor(t, dtype=dx.dtype).reshape(1, 1)
    for i, n in enumerate(names):
        p = params[n]
        dx = (dx @ p["w"] + p["b"]) * sigmoid(t @ p["w_gate"] + p["b_gate"]) + t @ p["w_bias"]
        if i < len(names) - 1:
            dx = softplus(dx)
    return dx


def css_taylor_closed(params, x, t):
    """Closed-form (f, d^2 z/dt^2) along z' = f(z, t), SURVEY.md Appendix B:
    G = sigmoid(w_g t + b_g), lin = x W + b, u = lin*G + w_b t;  G' = G(1-G) w_g;
    u' = (x' W)*G + lin*G' + w_b;  softplus' = sigmoid."""
    names = layer_names(len(params))
    f = css_dynamics(params, x, t)
    z = x.reshape(-1, params[names[0]]["w"].shape[0])
    z1 = f
    tt = torch.as_tensor(t, dtype=z.dtype).reshape(1, 1)
    for i, n in enumerate(names):
        p = params[n]
        gate = sigmoid(tt @ p["w_gate"] + p["b_gate"])
        gate1 = gate * (1 - gate) * p["w_gate"]
        lin = z @ p["w"] + p["b"]
        u = lin * gate + tt @ p["w_bias"]
        u1 = (z1 @ p["w"]) * gate + lin * gate1 + p["w_bias"]
        if i < len(names) - 1:
            z, z1 = softplus(u), sigmoid(u) * u1
        else:
            z, z1 = u, u1
    return f, z1


def make_ffjord_closures(reg="r2", reg_type="our"):
    """The closures init_model() builds (ffjord_tabular.py:235-297), over explicit params."""
    def dynamics_wrap(x, t, params):
        return css_dynamics(params, x, t)

    def reg_dynamics(y, t, params):
        if reg == "none":
            return torch.zeros((y.shape[0], 1), dtype=y.dtype)
        _, r = sol_recursive(lambda _y, _t: dynamics_wrap(_y, _t, params), y, t, "r2")   # K = 2 whatever --reg says
        return torch.mean(torch.square(r), dim=tuple(range(1, r.ndim)))

    def ffjord2_dynamics(yp, t, eps, params):
        y = yp[0]
        dy, eps_dy = torch.func.jvp(lambda _y: dynamics_wrap(_y, t, params), (y,), (eps,))
        div = torch.sum((eps_dy * eps).reshape(y.shape[0], -1), dim=1, keepdim=True)
        return dy, -div, eps_dy

    def ffjord_dynamics(yp, t, eps, params):
        return ffjord2_dynamics(yp, t, eps, params)[:2]

    def aug_dynamics(ypr, t, eps, params):
        y, p = ypr[0], ypr[1]
        dy, dp, eps_dy = ffjord2_dynamics((y, p), t, eps, params)
        if reg_type == "our":
            return dy, dp, reg_dynamics(y, t, params)
        dims = tuple(range(1, dy.ndim))
        return dy, dp, torch.mean(torch.square(eps_dy), dim=dims), torch.mean(torch.square(dy), dim=dims)

    def all_aug_dynamics(ypr, t, eps, params):
        y, p = ypr[0], ypr[1]
        dy, dp, eps_dy = ffjord2_dynamics((y, p), t, eps, params)
        dims = tuple(range(1, dy.ndim))
        return (dy, dp, reg_dynamics(y, t, params), torch.mean(torch.square(eps_dy), dim=dims),
                torch.mean(torch.square(dy), dim=dims))

    fwd = lambda y, t, eps, params: dynamics_wrap(y, t, params)
    return dict(dynamics_wrap=dynamics_wrap, reg_dynamics=reg_dynamics, ffjord_dynamics=ffjord_dynamics,
                ffjord2_dynamics=ffjord2_dynamics, aug_dynamics=aug_dynamics, all_aug_dynamics=all_aug_dynamics, fwd=fwd)


# ---- one update() of ffjord_tabular.py (the CPU arm of bench.py --workload ffjord_tabular) ----------------------------
def standard_normal_logprob(z):
    """ffjord_tabular.py:402-407."""
    return -0.5 * math.log(2 * math.pi) - torch.square(z) / 2


def train_step_grads(params, x, eps, lam=1e-2, lam_w=1e-6, reg="r2", rtol=1.4e-8, atol=1.4e-8):
    """jax.grad(loss_fn) of ffjord_tabular.py:417-432, 521 for reg_type 'our': forward_aux = odeint_sepaux (forward solves y
    only; delta_logp and r come back as zeros, lib/ode.py:1006-1018), loss = -mean(logpz - delta_logp) + lam mean(r) +
    lam_w 0.5 |theta|^2; the backward is the adjoint of aug_dynamics (lib/ode.py:1686-1693)."""
    from . import ode_oracle as oo
    cl = make_ffjord_closures(reg, "our")
    B = x.shape[0]
    ts = torch.tensor([0., 1.], dtype=x.dtype)
    z0 = torch.zeros(B, 1, dtype=x.dtype)
    res, nfe, fst = oo.odeint_split_fwd(cl["fwd"], (x, z0, z0), ts, eps, params, n_aux=2, rtol=rtol, atol=atol, return_stats=True)
    z = res[0][-1]
    logpz = torch.sum(standard_normal_logprob(z).reshape(B, -1), dim=1, keepdim=True)
    loss = -torch.mean(logpz - res[1][-1])
    gg = tuple(torch.zeros_like(r) for r in res)
    gg[0][-1], gg[1][-1], gg[2][-1] = z / B, 1.0 / B, lam / B
    bst = []
    out = oo.odeint_split_rev(cl["aug_dynamics"], rtol, atol, math.inf, res, ts, (eps, params), gg, bst)
    grads = {n: {k: out[3][n][k] + lam_w * params[n][k] for k in params[n]} for n in params}
    return dict(loss=loss, nfe=nfe, grads=grads, fwd_stats=fst, bwd_stats=bst, z=z)


def adam_init(params):
    zeros = lambda: {n: {k: torch.zeros_like(v) for k, v in p.items()} for n, p in params.items()}
    return zeros(), zeros()


def adam_update(params, m, v, grads, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """jax.experimental.optimizers.adam (ffjord_tabular.py:527)."""
    new_p, new_m, new_v = {}, {}, {}
    for n in params:
        new_p[n], new_m[n], new_v[n] = {}, {}, {}
        for k in params[n]:
            g = grads[n][k]
            mi = (1 - b1) * g + b1 * m[n][k]
            vi = (1 - b2) * g * g + b2 * v[n][k]
            mhat = mi / (1 - b1 ** (step + 1))
            vhat = vi / (1 - b2 ** (step + 1))
            new_p[n][k] = params[n][k] - lr * mhat / (torch.sqrt(vhat) + eps)
            new_m[n][k], new_v[n][k] = mi, vi
    return new_p, new_m, new_v
