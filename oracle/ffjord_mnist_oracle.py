"""CPU oracle for the closures of ffjord_mnist.py (BASELINE config C4): ConcatConv2D dynamics, the FFJORD closures on an
image state, the logit pre-flow and the bits/dim loss.

TEST INFRASTRUCTURE.  Restates, in eager torch:
  _logaddexp / softplus ... ffjord_mnist.py:85-96
  ConcatConv2D ............ ffjord_mnist.py:123-135   hk.Conv2D on concat([t * ones, x], channel axis); NHWC, 3x3, pad 1
  NN_Dynamics ............. ffjord_mnist.py:186-232   hidden_dims (64, 64, 64) + (1,), strides 1, softplus between
                                                      layers, none after the last, LAST LAYER'S w ZERO-INITIALISED
  sol_recursive ........... ffjord_mnist.py:99-119    fixed K = 2 whatever --reg says
  reg_dynamics ............ ffjord_mnist.py:278-289
  ffjord(2)_dynamics ...... ffjord_mnist.py:291-309   jvp with eps (Rademacher, 147-152): div = sum eps * (J eps)
  aug_dynamics ............ ffjord_mnist.py:311-324   (dy, -div, r) or (dy, -div, fro, kin)
  all_aug_dynamics ........ ffjord_mnist.py:326-338
  ForwardPreODE ........... ffjord_mnist.py:155-168   logit pre-flow, logdetgrad 138-144
  _loss_fn ................ ffjord_mnist.py:455-470   bits per dim

PARITY UNPINNED for the Taylor-mode part (jax.experimental.jet, see dynamics_oracle.py); hk.Conv2D is restated as
torch.nn.functional.conv2d (cross-correlation, like lax.conv_general_dilated) with the HWIO weight permuted to OIHW.
Nothing under easy-neural-ode_b200/ imports this module.
"""
import math

import torch
import torch.nn.functional as F

from .dynamics_oracle import sol_recursive
from .ffjord_oracle import softplus


def layer_names(n_layers):
    """Haiku module names of NN_Dynamics' ConcatConv2D stack, in call order (= sorted order: ravel_pytree order)."""
    return ["nn__dynamics/concat_conv2_d" + ("" if i == 0 else f"_{i}") + "/conv2_d" for i in range(n_layers)]


def init_conv_params(hidden=64, n_hidden_layers=3, channels=1, seed=0, dtype=torch.float32, scale=1.0, last_scale=0.0):
    """hk.Conv2D defaults: w [3, 3, c_in + 1, c_out] ~ truncated normal, sigma = 1 / sqrt(fan_in), fan_in = 9 (c_in + 1);
    b [c_out] zeros.  The script zero-initialises the last layer's w (ffjord_mnist.py:214-215): ``last_scale`` = 0
    reproduces that (f == 0 at initialisation), a non-zero value gives the trained-like dynamics the tests need."""
    g = torch.Generator().manual_seed(seed)

    def tn(shape, fan_in, s):
        w = torch.empty(shape, dtype=torch.float64)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
        return (w * (s / math.sqrt(fan_in))).to(dtype)

    dims = [channels] + [hidden] * n_hidden_layers + [channels]
    names = layer_names(len(dims) - 1)
    params = {}
    for i, (name, (c_in, c_out)) in enumerate(zip(names, zip(dims[:-1], dims[1:]))):
        last = i == len(names) - 1
        params[name] = {"b": torch.zeros(c_out, dtype=dtype),
                        "w": tn((3, 3, c_in + 1, c_out), 9 * (c_in + 1), last_scale if last else scale)}
    return params


def conv_dynamics(params, x, t, image=(28, 28, 1)):
    """f(x, t) of ffjord_mnist.py:219-232; x: anything reshapeable to [B, H, W, C] (NHWC as in the script)."""
    names = layer_names(len(params))
    dx = x.reshape(-1, *image)
    t = torch.as_tensor(t, dtype=dx.dtype)
    for i, n in enumerate(names):
        p = params[n]
        ttx = torch.cat([torch.ones_like(dx[..., :1]) * t, dx], dim=-1)          # time is input channel 0
        y = F.conv2d(ttx.permute(0, 3, 1, 2).contiguous(), p["w"].permute(3, 2, 0, 1).contiguous(), bias=p["b"], stride=1, padding=1)
        dx = y.permute(0, 2, 3, 1)
        if i < len(names) - 1:
            dx = softplus(dx)
    return dx


def conv_taylor_closed(params, x, t, image=(28, 28, 1)):
    """Closed-form (f, d^2 z/dt^2) along z' = f(z, t): one tangent traversal with direction (f, dt = 1); the time
    channel's tangent is a constant 1 inside the zero padding."""
    names = layer_names(len(params))
    f = conv_dynamics(params, x, t, image)
    z = x.reshape(-1, *image)
    z1 = f
    t = torch.as_tensor(t, dtype=z.dtype)
    for i, n in enumerate(names):
        p = params[n]
        w = p["w"].permute(3, 2, 0, 1).contiguous()
        ttx = torch.cat([torch.ones_like(z[..., :1]) * t, z], dim=-1).permute(0, 3, 1, 2)
        tt1 = torch.cat([torch.ones_like(z[..., :1]), z1], dim=-1).permute(0, 3, 1, 2)
        u = F.conv2d(ttx, w, bias=p["b"], padding=1).permute(0, 2, 3, 1)
        u1 = F.conv2d(tt1, w, bias=None, padding=1).permute(0, 2, 3, 1)
        if i < len(names) - 1:
            z, z1 = softplus(u), torch.sigmoid(u) * u1
        else:
            z, z1 = u, u1
    return f, z1


def make_ffjord_mnist_closures(reg="r2", reg_type="our", image=(28, 28, 1)):
    """The closures init_model() builds (ffjord_mnist.py:268-338), over explicit params; states are [B, H, W, C]."""
    def dynamics_wrap(x, t, params):
        return conv_dynamics(params, x, t, image)

    def reg_dynamics(y, t, params):
        y = y.reshape(-1, *image)
        if reg == "none":
            return torch.zeros((y.shape[0], 1), dtype=y.dtype)
        _, r = sol_recursive(lambda _y, _t: dynamics_wrap(_y, _t, params), y, t, "r2")   # K = 2 whatever --reg says
        return torch.mean(torch.square(r), dim=tuple(range(1, r.ndim)))

    def ffjord2_dynamics(yp, t, eps, params):
        y = yp[0].reshape(-1, *image)
        dy, eps_dy = torch.func.jvp(lambda _y: dynamics_wrap(_y, t, params), (y,), (eps.reshape(y.shape),))
        div = torch.sum((eps_dy * eps.reshape(y.shape)).reshape(y.shape[0], -1), dim=1, keepdim=True)
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


# ---- glue around the ODE block ------------------------------------------------------------------------------------
def logdetgrad(x, alpha):
    """ffjord_mnist.py:138-144."""
    s = alpha + (1 - 2 * alpha) * x
    ld = -torch.log(s - s * s) + math.log(1 - 2 * alpha)
    return torch.sum(ld.reshape(x.shape[0], -1), dim=1, keepdim=True)


def forward_pre_ode(x, logpx, alpha=1e-5):
    """ForwardPreODE, ffjord_mnist.py:155-168: y = logit(alpha + (1 - 2 alpha) x), logpy = logpx - logdetgrad(x)."""
    s = alpha + (1 - 2 * alpha) * x
    return torch.log(s) - torch.log(1 - s), logpx - logdetgrad(x, alpha)


def standard_normal_logprob(z):
    return -0.5 * math.log(2 * math.pi) - torch.square(z) / 2


def bits_per_dim(z, delta_logp):
    """_loss_fn, ffjord_mnist.py:455-470."""
    logpz = torch.sum(standard_normal_logprob(z).reshape(z.shape[0], -1), dim=1, keepdim=True)
    logpx = logpz - delta_logp
    logpx_per_dim = torch.sum(logpx) / z.numel()
    return -(logpx_per_dim - math.log(256)) / math.log(2)


# ---- one update() of ffjord_mnist.py (the CPU arm of bench.py --workload ffjord_mnist) ---------------------------------
def train_step_grads(params, images, eps, lam=3e-4, lam_w=0.0, reg="r2", rtol=1.4e-8, atol=1.4e-8, image=(28, 28, 1)):
    """jax.grad(loss_fn) of ffjord_mnist.py:480-497, 544 for reg_type 'our': forward_aux = pre-flow + odeint_sepaux (forward
    solves y only; delta_logp and r come back as zeros, lib/ode.py:1006-1018), loss = bits/dim + lam mean(r) + lam_w 0.5
    |theta|^2; the backward is the adjoint of aug_dynamics (lib/ode.py:1686-1693)."""
    from . import ode_oracle as oo
    cl = make_ffjord_mnist_closures(reg, "our", image)
    B = images.shape[0]
    ts = torch.tensor([0., 1.], dtype=images.dtype)
    z0 = torch.zeros(B, 1, dtype=images.dtype)
    y, logpy = forward_pre_ode(images.reshape(-1, *image), z0)
    res, nfe, fst = oo.odeint_split_fwd(cl["fwd"], (y, logpy, z0), ts, eps, params, n_aux=2, rtol=rtol, atol=atol, return_stats=True)
    z = res[0][-1]
    loss = bits_per_dim(z, res[1][-1])
    gg = tuple(torch.zeros_like(r) for r in res)
    nd = z.numel() * math.log(2)
    gg[0][-1], gg[1][-1], gg[2][-1] = z / nd, 1.0 / nd, lam / B
    bst = []
    out = oo.odeint_split_rev(cl["aug_dynamics"], rtol, atol, math.inf, res, ts, (eps, params), gg, bst)
    grads = {n: {k: out[3][n][k] + lam_w * params[n][k] for k in params[n]} for n in params}
    return dict(loss=loss, nfe=nfe, grads=grads, fwd_stats=fst, bwd_stats=bst, z=z)
