"""CPU oracle for the script-level closures around the solver (mnist.py).

TEST INFRASTRUCTURE.  Restates, in eager torch:
  sigmoid ................. mnist.py:90-94
  MLPDynamics ............. mnist.py:195-222   (time is column 0 of each concat)
  sol_recursive ........... mnist.py:105-131   (jet recursion, see below)
  reg_dynamics ............ mnist.py:287-292
  fin_dynamics ............ mnist.py:294-300
  aug_dynamics ............ mnist.py:302-313
  all_aug_dynamics ........ mnist.py:315-324
  aug_init / all_aug_init . mnist.py:416-438
  PostODE, losses ......... mnist.py:225-238, 449-478
  momentum update ......... mnist.py:530-540 (jax.experimental.optimizers.momentum)

PARITY UNPINNED for the Taylor-mode part: the reference calls
``jax.experimental.jet`` (JAX, version never pinned by the repo; call site
mnist.py:127-129), which is absent from /root/reference and not installable
here.  jet propagates exact derivative coefficients d^k/dtau^k (un-normalised:
the recursion feeds outputs straight back as inputs), so it is restated here by
its published definition: ``z^(1) = g(z)``, ``z^(k+1) = d/dtau z^(k)`` along the
autonomous lift ``g([z,t]) = [f(z,t), 1]``, computed with nested forward-mode
``torch.func.jvp``.  An independent closed form for the MLP (``mlp_taylor_closed``)
and a finite-difference check of the true ODE flow cross-check it in
tests/test_oracle_dynamics.py.
"""
import math

import torch

L1 = "mlp_dynamics/linear"
L2 = "mlp_dynamics/linear_1"
REGS = ["r2", "r3", "r4", "r5", "r6"]


def sigmoid(z):
    return 1 / (1 + torch.exp(-z))


def init_mlp_params(dim, hidden, seed=0, dtype=torch.float32, scale=1.0):
    """Haiku-default Linear init: w ~ truncated normal(+-2 sigma), sigma = 1/sqrt(fan_in),
    b = 0.  ``scale`` multiplies the weights to emulate trained (stiffer) dynamics."""
    g = torch.Generator().manual_seed(seed)

    def tn(shape, fan_in):
        w = torch.empty(shape, dtype=torch.float64)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
        return (w * (scale / math.sqrt(fan_in))).to(dtype)

    return {
        L1: {"w": tn((dim + 1, hidden), dim + 1), "b": torch.zeros(hidden, dtype=dtype)},
        L2: {"w": tn((hidden + 1, dim), hidden + 1), "b": torch.zeros(dim, dtype=dtype)},
    }


def mlp_dynamics(params, x, t):
    """f(x, t) of mnist.py:208-222; x: [B, D] (or [D]) -> [B, D]."""
    w1, b1 = params[L1]["w"], params[L1]["b"]
    w2, b2 = params[L2]["w"], params[L2]["b"]
    x = x.reshape(-1, w1.shape[0] - 1)
    out = sigmoid(x)
    tt = torch.ones_like(x[:, :1]) * t
    out = torch.cat([tt, out], dim=-1) @ w1 + b1
    out = sigmoid(out)
    tt = torch.ones_like(out[:, :1]) * t
    out = torch.cat([tt, out], dim=-1) @ w2 + b2
    return out


def sol_recursive(f, z, t, reg):
    """-> (dz/dt, d^K z/dt^K), K = 2 for r2 ... 6 for r6; zeros for 'none'."""
    if reg == "none":
        return f(z, t), torch.zeros_like(z)
    K = REGS.index(reg) + 2
    shape = z.shape
    t = torch.as_tensor(t, dtype=z.dtype)
    zt = torch.cat([z.reshape(-1), t.reshape(1)])

    def g(v):
        dz = f(v[:-1].reshape(shape), v[-1]).reshape(-1)
        return torch.cat([dz, torch.ones(1, dtype=v.dtype)])

    def deriv(k):
        if k == 1:
            return g
        prev = deriv(k - 1)
        return lambda v: torch.func.jvp(prev, (v,), (g(v),))[1]

    return g(zt)[:-1].reshape(shape), deriv(K)(zt)[:-1].reshape(shape)


def mlp_taylor_closed(params, z, t, order):
    """Closed-form z^(1..order) (order <= 3) for MLPDynamics, SURVEY.md Appendix B."""
    w1, b1 = params[L1]["w"], params[L1]["b"]
    w2, b2 = params[L2]["w"], params[L2]["b"]
    w1t, w1x, w2t, w2x = w1[0], w1[1:], w2[0], w2[1:]
    s = sigmoid(z)
    a = t * w1t + s @ w1x + b1
    h = sigmoid(a)
    f = t * w2t + h @ w2x + b2
    out = [f]
    if order >= 2:
        sp, hp = s * (1 - s), h * (1 - h)
        s1 = sp * f
        a1 = w1t + s1 @ w1x
        h1 = hp * a1
        y1 = w2t + h1 @ w2x
        out.append(y1)
    if order >= 3:
        spp, hpp = sp * (1 - 2 * s), hp * (1 - 2 * h)
        s2 = spp * f * f + sp * y1
        a2 = s2 @ w1x
        h2 = hpp * a1 * a1 + hp * a2
        y2 = h2 @ w2x
        out.append(y2)
    return out


def mlp_dynamics_xtp(x, t, params):
    """dynamics_wrap of mnist.py:285 (argument order of the solver: state, time, *args)."""
    return mlp_dynamics(params, x, t)


def make_mnist_closures(reg="r3", reg_type="our"):
    """The closures init_model() builds (mnist.py:285-334), over explicit params."""
    def dynamics_wrap(x, t, params):
        return mlp_dynamics(params, x, t)

    def reg_dynamics(y, t, params):
        y0, r = sol_recursive(lambda _y, _t: dynamics_wrap(_y, _t, params), y, t, reg)
        return y0, torch.mean(torch.square(r), dim=tuple(range(1, r.ndim)))

    def fin_dynamics(y, t, eps, params):
        return torch.func.jvp(lambda _y: dynamics_wrap(_y, t, params), (y,), (eps,))

    def aug_dynamics(yr, t, eps, params):
        y = yr[0]
        if reg_type == "our":
            return reg_dynamics(y, t, params)
        dy, eps_dy = fin_dynamics(y, t, eps, params)
        dims = tuple(range(1, dy.ndim))
        return dy, torch.mean(torch.square(eps_dy), dim=dims), torch.mean(torch.square(dy), dim=dims)

    def all_aug_dynamics(yr, t, eps, params):
        y = yr[0]
        dy, eps_dy = fin_dynamics(y, t, eps, params)
        _, drdt = reg_dynamics(y, t, params)
        dims = tuple(range(1, dy.ndim))
        return dy, drdt, torch.mean(torch.square(eps_dy), dim=dims), torch.mean(torch.square(dy), dim=dims)

    fwd = lambda y, t, eps, params: dynamics_wrap(y, t, params)
    return dict(dynamics_wrap=dynamics_wrap, reg_dynamics=reg_dynamics, fin_dynamics=fin_dynamics,
                aug_dynamics=aug_dynamics, all_aug_dynamics=all_aug_dynamics, fwd=fwd)


# ----------------------------------------------------------------------------
# one training step of mnist.py (reg_type == 'our'): forward solve, adjoint,
# momentum update.  Used for gradient parity and as the CPU baseline.
# ----------------------------------------------------------------------------
def init_post_params(dim, n_cls=10, seed=1, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    w = torch.empty((dim, n_cls), dtype=torch.float64)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
    return {"w": (w / math.sqrt(dim)).to(dtype), "b": torch.zeros(n_cls, dtype=dtype)}


def pre_ode(images_u8, dtype=torch.float32):
    return (images_u8.to(dtype) / 255.).reshape(images_u8.shape[0], -1)


def head_loss_and_grad(post, y1, labels):
    """PostODE (sigmoid -> Linear(10)) + mean softmax CE; returns loss, dL/dy1, dL/dpost."""
    y1 = y1.detach().requires_grad_(True)
    w = post["w"].detach().requires_grad_(True)
    b = post["b"].detach().requires_grad_(True)
    logits = sigmoid(y1) @ w + b
    loss = -(torch.log_softmax(logits, dim=-1)[torch.arange(labels.shape[0]), labels]).mean()
    gy, gw, gb = torch.autograd.grad(loss, (y1, w, b))
    return loss.detach(), gy, {"w": gw, "b": gb}


def train_step_grads(ode_params, post, x0, labels, eps, lam, reg="r3", rtol=1.4e-8, atol=1.4e-8,
                     collect=None, reg_type="our", lam_fro=0.0, lam_kin=0.0):
    """loss_fn + jax.grad of mnist.py:462-478, 528, lam_w = 0; reg_type 'our' (odeint_aux_one, lam * mean r) or
    'fin' (odeint_sepaux, lam_fro * mean fro + lam_kin * mean kin).

    -> dict(loss, y1, nfe, grads={'ode':..., 'post':...}, x0_bar, fwd_stats, bwd_stats)
    """
    from . import ode_oracle as oo
    cl = make_mnist_closures(reg, reg_type)
    B = x0.shape[0]
    ts = torch.tensor([0., 1.], dtype=x0.dtype)
    if reg_type == "fin":
        zero = torch.zeros(B, dtype=x0.dtype)
        (ys, fs, ks), nfe, fst = oo.odeint_split_fwd(cl["fwd"], (x0, zero, zero.clone()), ts, eps, ode_params, n_aux=2,
                                                     rtol=rtol, atol=atol, return_stats=True)
        loss, gy, gpost = head_loss_and_grad(post, ys[-1], labels)
        g_y, g_f, g_k = torch.zeros_like(ys), torch.zeros_like(fs), torch.zeros_like(ks)
        g_y[-1], g_f[-1], g_k[-1] = gy, lam_fro / B, lam_kin / B
        bst = []
        y0_bar, ts_bar, eps_bar, params_bar = oo.odeint_split_rev(
            cl["aug_dynamics"], rtol, atol, math.inf, (ys, fs, ks), ts, (eps, ode_params), (g_y, g_f, g_k), bst)
        return dict(loss=loss, y1=ys[-1], nfe=nfe, grads={"ode": params_bar, "post": gpost},
                    x0_bar=y0_bar[0], ts_bar=ts_bar, eps_bar=eps_bar, fwd_stats=fst, bwd_stats=bst)
    y0 = (x0, torch.zeros(B, dtype=x0.dtype))
    (ys, rs), nfe, fst = oo.odeint_split_fwd(cl["fwd"], y0, ts, eps, ode_params, n_aux=1,
                                             rtol=rtol, atol=atol, return_stats=True)
    loss, gy, gpost = head_loss_and_grad(post, ys[-1], labels)
    g_y = torch.zeros_like(ys)
    g_y[-1] = gy
    g_r = torch.zeros_like(rs)
    g_r[-1] = lam / B
    bst = []
    y0_bar, ts_bar, eps_bar, params_bar = oo.odeint_split_rev(
        cl["aug_dynamics"], rtol, atol, math.inf, (ys, rs), ts, (eps, ode_params), (g_y, g_r), bst)
    return dict(loss=loss, y1=ys[-1], nfe=nfe, grads={"ode": params_bar, "post": gpost},
                x0_bar=y0_bar[0], ts_bar=ts_bar, fwd_stats=fst, bwd_stats=bst)


def ode_oracle_tree_zeros(tree):
    return {k: (ode_oracle_tree_zeros(v) if isinstance(v, dict) else torch.zeros_like(v)) for k, v in tree.items()}


def momentum_update(params, velocity, grads, lr, mass=0.9):
    """jax.experimental.optimizers.momentum: v <- mass*v + g ; x <- x - lr*v."""
    new_p, new_v = {}, {}
    for k in params:
        if isinstance(params[k], dict):
            new_p[k], new_v[k] = momentum_update(params[k], velocity[k], grads[k], lr, mass)
        else:
            new_v[k] = mass * velocity[k] + grads[k]
            new_p[k] = params[k] - lr * new_v[k]
    return new_p, new_v
