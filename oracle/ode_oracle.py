"""CPU oracle: restatement of easy-neural-ode's adaptive explicit-RK odeint.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product path
(easy-neural-ode_b200/) never does and fails loudly without its CUDA library.

Parity status: the SOLVER CORE (this file) is PINNED - tests/golden/*.npz were
produced by running the unmodified /root/reference/lib/ode.py on the torch-backed
JAX stand-in (oracle/jax_shim.py, script tests/golden/make_golden.py) and this
restatement reproduces them.  What stays UNPINNED is the third-party arithmetic
the reference delegates to JAX (jax.experimental.jet, XLA's reduction order):
see oracle/dynamics_oracle.py.

Follows (file:line under /root/reference/):
  initial_step_size ........ lib/ode.py:86-104
  rk_step .................. lib/ode.py:106-134 and the 8 sibling steps 136-378
  fit_quartic / interp ..... lib/ode.py:56-63, 78-84, 849-850
  error_ratio, next_dt ..... lib/ode.py:518-538
  solve (dopri5 flavour) ... lib/ode.py:823-862
  solve (clip flavour) ..... lib/ode.py:1048-1390
  odeint + ravel ........... lib/ode.py:46-54, 540-559, 742-747
  adjoint .................. lib/ode.py:1442-1444, 1494-1529
  split wrappers ........... lib/ode.py:993-1032, 1478-1488, 1677-1693

Everything is eager torch on CPU; ``dtype`` of the state decides fp32 / fp64
(scalars t, dt, ratio live in that dtype too, like the reference's traced
scalars).
"""
import math

import torch

from . import tableaux as _tb


# ----------------------------------------------------------------------------
# pytree helpers (tuple / list / dict in sorted-key order), cf. ravel_pytree
# ----------------------------------------------------------------------------
def _flatten(tree):
    if isinstance(tree, dict):
        ks = sorted(tree)
        parts = [_flatten(tree[k]) for k in ks]
        return [l for p in parts for l in p[0]], ("d", ks, [p[1] for p in parts])
    if isinstance(tree, (tuple, list)):
        parts = [_flatten(v) for v in tree]
        return [l for p in parts for l in p[0]], ("t" if isinstance(tree, tuple) else "l", None, [p[1] for p in parts])
    return [tree], None


def _count(td):
    return 1 if td is None else sum(_count(c) for c in td[2])


def _unflatten(td, leaves):
    if td is None:
        return leaves[0]
    out, pos = [], 0
    for c in td[2]:
        n = _count(c)
        out.append(_unflatten(c, leaves[pos:pos + n]))
        pos += n
    if td[0] == "d":
        return dict(zip(td[1], out))
    return tuple(out) if td[0] == "t" else out


def tree_map(fn, tree):
    leaves, td = _flatten(tree)
    return _unflatten(td, [fn(l) for l in leaves])


def ravel(tree, dtype=None):
    """-> (flat vector, unravel).  Python scalars become 0-d tensors."""
    leaves, td = _flatten(tree)
    leaves = [l if isinstance(l, torch.Tensor) else torch.as_tensor(l, dtype=dtype or torch.get_default_dtype())
              for l in leaves]
    shapes = [l.shape for l in leaves]
    sizes = [l.numel() for l in leaves]
    flat = torch.cat([l.reshape(-1) for l in leaves]) if leaves else torch.zeros(0)

    def unravel(v):
        out, pos = [], 0
        for sh, n in zip(shapes, sizes):
            out.append(v[pos:pos + n].reshape(sh))
            pos += n
        return _unflatten(td, out)

    return flat, unravel


# ----------------------------------------------------------------------------
# solver core
# ----------------------------------------------------------------------------
def _tab(name, dtype):
    np_dt = {torch.float32: "float32", torch.float64: "float64"}[dtype]
    t = _tb.as_dtype(name, np_dt)
    out = dict(t)
    for k in ("alpha", "beta", "c_sol", "c_err", "c_mid"):
        if k in t:
            out[k] = torch.from_numpy(t[k])
    return out


def initial_step_size(fun, t0, y0, order, rtol, atol, f0):
    """Hairer II.4 starting step with Euclidean norms (lib/ode.py:86-104)."""
    scale = atol + torch.abs(y0) * rtol
    d0 = torch.linalg.norm(y0 / scale)
    d1 = torch.linalg.norm(f0 / scale)
    if bool((d0 < 1e-5) | (d1 < 1e-5)):
        h0 = torch.as_tensor(1e-6, dtype=y0.dtype)
    else:
        h0 = 0.01 * d0 / d1
    f1 = fun(y0 + h0 * f0, t0 + h0)
    d2 = torch.linalg.norm((f1 - f0) / scale) / h0
    if bool((d1 <= 1e-15) & (d2 <= 1e-15)):
        h1 = torch.maximum(torch.as_tensor(1e-6, dtype=y0.dtype), h0 * 1e-3)
    else:
        h1 = (0.01 / (d1 + d2)) ** (1. / (order + 1.))
    return torch.minimum(100. * h0, h1)


def rk_step(fun, y0, f0, t0, dt, tab):
    """One embedded step: stage derivatives, 5th/4th.. solutions, FSAL-style f1.

    ``f1 = k[-1]`` for every tableau, FSAL or not (lib/ode.py:133, 158, ... 377).
    """
    s = tab["stages"]
    k = torch.zeros((s, f0.shape[0]), dtype=y0.dtype)
    k[0] = f0
    for i in range(1, s):
        ti = t0 + dt * tab["alpha"][i - 1]
        yi = y0 + dt * torch.matmul(tab["beta"][i - 1], k)
        k = k.clone()
        k[i] = fun(yi, ti)
    y1 = dt * torch.matmul(tab["c_sol"], k) + y0
    err = dt * torch.matmul(tab["c_err"], k)
    return y1, k[-1], err, k


def fit_quartic(y0, y1, y_mid, dy0, dy1, dt):
    """Quartic through (y0, y_mid, y1) with end slopes (lib/ode.py:78-84)."""
    a = -2. * dt * dy0 + 2. * dt * dy1 - 8. * y0 - 8. * y1 + 16. * y_mid
    b = 5. * dt * dy0 - 3. * dt * dy1 + 18. * y0 + 14. * y1 - 32. * y_mid
    c = -4. * dt * dy0 + dt * dy1 - 11. * y0 - 5. * y1 + 16. * y_mid
    d = dt * dy0
    e = y0
    return torch.stack([a, b, c, d, e])


def polyval(coeffs, x):
    y = torch.zeros_like(coeffs[0] * x)
    for c in coeffs:
        y = y * x + c
    return y


def error_ratio(err, rtol, atol, y0, y1):
    """mean((err / (atol + rtol*max(|y0|,|y1|)))^2) over the whole flat state."""
    tol = atol + rtol * torch.maximum(torch.abs(y0), torch.abs(y1))
    return torch.mean(torch.square(err / tol))


def next_dt(last, ratio, order=5.0, safety=0.9, ifactor=10.0, dfactor=0.2):
    """Step-size controller (lib/ode.py:529-538)."""
    if bool(ratio < 1):
        dfactor = 1.0
    factor = torch.clamp(torch.sqrt(ratio) ** (1.0 / order) / safety, max=1.0 / dfactor)
    factor = torch.clamp(factor, min=1.0 / ifactor)
    if bool(ratio == 0):
        return last * ifactor
    return last / factor


def solve(fun, y0, ts, rtol, atol, mxstep=math.inf, solver="dopri5", trace=None):
    """Adaptive solve on a flat state.  -> (ys [T,N], nfe, stats).

    dopri5: no dt clipping, overshoot + quartic dense output from the c_mid
    combination (lib/ode.py:823-862).  Every other tableau: dt clipped to the
    target, mid-point from a second step of dt/2 whose evaluations are not
    counted, and the interpolant anchored at the SOLVE's initial y0
    (lib/ode.py:1048-1390).
    """
    dtype = y0.dtype
    tab = _tab(solver, dtype)
    ts = ts.to(dtype)
    rtol = float(rtol)
    atol = float(atol)
    y_init = y0
    f = fun(y0, ts[0])
    dt = initial_step_size(fun, ts[0], y0, tab["init_order"], rtol, atol, f)
    nfe = 2.0
    y, t, last_t = y0, ts[0], ts[0]
    coeff = torch.stack([y0] * 5)
    outs = [y0]
    n_iters = n_acc = 0
    for target in ts[1:]:
        i = 0
        while bool(t < target) and i < mxstep and bool(dt > 0):
            if tab["clip"]:
                if bool(t + dt > target):
                    dt = target - t
            y1, f1, err, k = rk_step(fun, y, f, t, dt, tab)
            t1 = t + dt
            ratio = error_ratio(err, rtol, atol, y, y1)
            if tab["clip"]:
                y_mid = rk_step(fun, y, f, t, dt / 2, tab)[0]
                new_coeff = fit_quartic(y_init, y1, y_mid, k[0], k[-1], dt)
            else:
                y_mid = y + dt * torch.matmul(tab["c_mid"], k)
                new_coeff = fit_quartic(y, y1, y_mid, k[0], k[-1], dt)
            dt_new = next_dt(dt, ratio, order=float(tab["order"]))
            ok = bool(ratio <= 1.)
            if trace is not None:
                trace.append((float(t), float(dt), float(ratio), ok))
            if ok:
                y, f, last_t, t, coeff = y1, f1, t, t1, new_coeff
                n_acc += 1
            dt = dt_new
            i += 1
        n_iters += i
        nfe += tab["nfe_per_iter"] * i
        rel = (target - last_t) / (t - last_t)
        outs.append(polyval(coeff, rel))
    stats = dict(n_iters=n_iters, n_accepted=n_acc, last_dt=float(dt))
    return torch.stack(outs), nfe, stats


# ----------------------------------------------------------------------------
# public surface (mirrors lib/ode.py:540-559, 657-718)
# ----------------------------------------------------------------------------
def odeint(func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=math.inf, solver="dopri5",
           return_stats=False, trace=None):
    flat0, unravel = ravel(y0)

    def fun(yflat, tt):
        return ravel(func(unravel(yflat), tt, *args))[0]

    ys, nfe, stats = solve(fun, flat0, t, rtol, atol, mxstep, solver, trace)
    leaves_t = [unravel(ys[i]) for i in range(ys.shape[0])]
    l0, td = _flatten(leaves_t[0])
    stacked = [torch.stack([_flatten(lt)[0][j] for lt in leaves_t]) for j in range(len(l0))]
    out = _unflatten(td, stacked)
    if return_stats:
        return out, nfe, stats
    return out, nfe


def odeint_rev(func, rtol, atol, mxstep, ys, ts, args, g, stats=None):
    """Adjoint of ``odeint`` on a FLAT state (lib/ode.py:1494-1529).

    ys, g: [T, N]; args: tuple of pytrees.  Returns (y0_bar, ts_bar, *args_bar).
    Each segment is a fresh odeint over [-ts[i], -ts[i-1]] of the raveled tuple
    (y, y_bar, t0_bar, args_bar).
    """
    def aug(state, tneg, *a):
        y, y_bar = state[0], state[1]
        y_dot, pull = torch.func.vjp(func, y, -tneg, *a)
        return (-y_dot, *pull(y_bar))

    y_bar = g[-1]
    t0_bar = torch.zeros((), dtype=ys.dtype)
    args_bar = tree_map(torch.zeros_like, tuple(args))
    rev_ts_bar = []
    for i in range(len(ts) - 1, 0, -1):
        t_bar = torch.dot(func(ys[i], ts[i], *args), g[i])
        t0_bar = t0_bar - t_bar
        tr = [] if stats is not None else None
        out, _, st = odeint(aug, (ys[i], y_bar, t0_bar, args_bar),
                            torch.stack([-ts[i], -ts[i - 1]]), *args,
                            rtol=rtol, atol=atol, mxstep=mxstep, return_stats=True, trace=tr)
        if stats is not None:
            st = dict(st, trace=tr)     # controller trajectory [(t, dt, ratio, accepted)] of this segment
            stats.append(st)
        _, y_bar, t0_bar, args_bar = out
        y_bar, t0_bar, args_bar = tree_map(lambda x: x[1], (y_bar, t0_bar, args_bar))
        y_bar = y_bar + g[i - 1]
        rev_ts_bar.append(t_bar)
    ts_bar = torch.stack([t0_bar] + rev_ts_bar[::-1])
    return (y_bar, ts_bar, *args_bar)


def odeint_split_fwd(fwd_func, y0, t, *args, n_aux=1, rtol=1.4e-8, atol=1.4e-8, mxstep=math.inf,
                     return_stats=False):
    """Forward half of odeint_aux_one / odeint_sepaux / odeint_fin_sepaux
    (n_aux = 1 / 2 / 3): integrate ``y0[0]`` only with ``fwd_func`` and pad every
    auxiliary output with zeros of shape [T, B, 1] (lib/ode.py:993-1032)."""
    out, nfe, st = odeint(fwd_func, y0[0], t, *args, rtol=rtol, atol=atol, mxstep=mxstep, return_stats=True)
    T, B = out.shape[0], out.shape[1]
    res = (out,) + tuple(torch.zeros((T, B, 1), dtype=out.dtype) for _ in range(n_aux))
    return (res, nfe, st) if return_stats else (res, nfe)


def odeint_split_rev(rev_func, rtol, atol, mxstep, res_ys, ts, args, g, stats=None):
    """Backward half shared by the split wrappers (lib/ode.py:1677-1693):
    flatten (ys, zeros...) and g per time, run the plain adjoint with
    ``rev_func`` on the augmented flat state, un-flatten the y0 cotangent."""
    T = res_ys[0].shape[0]
    flat_res = torch.stack([ravel(tuple(r[i] for r in res_ys))[0] for i in range(T)])
    flat_g = torch.stack([ravel(tuple(x[i] for x in g))[0] for i in range(T)])
    _, unravel = ravel(tuple(r[0] for r in res_ys))

    def flat_func(yflat, tt, *a):
        return ravel(rev_func(unravel(yflat), tt, *a))[0]

    out = odeint_rev(flat_func, rtol, atol, mxstep, flat_res, ts, args, flat_g, stats)
    return (unravel(out[0]), *out[1:])


# ----------------------------------------------------------------------------
# fixed-grid RK4 family (lib/ode.py:561-634 wrappers, 864-946 primal, 1446-1460 fwd rules, 1531-1593 adjoint).
# PINNED by tests/golden/css_grid.npz (tests/golden/make_grid.py runs the unmodified reference on the JAX stand-in).
# ----------------------------------------------------------------------------
def rk4_solve(fun, step_size, y0, ts):
    """_rk4_odeint, lib/ode.py:864-903, on a flat state: steps of ``step_size`` from ts[0] while t < ts[-1], the last one
    clipped (next_t = min(t + step_size, ts[-1])); returns (stack(y0, y(ts[-1])), nfe = 4 per step)."""
    step_size = torch.as_tensor(step_size, dtype=y0.dtype)
    cur_y, cur_t, nfe = y0, ts[0], 0
    while bool(cur_t < ts[-1]):
        next_t = torch.minimum(cur_t + step_size, ts[-1])
        dt = next_t - cur_t
        k1 = fun(cur_y, cur_t)
        k2 = fun(cur_y + dt * k1 / 2, cur_t + dt / 2)
        k3 = fun(cur_y + dt * k2 / 2, cur_t + dt / 2)
        k4 = fun(cur_y + dt * k3, cur_t + dt)
        cur_y = cur_y + (k1 + 2 * k2 + 2 * k3 + k4) * (dt / 6)
        cur_t = next_t
        nfe += 4
    return torch.stack([y0, cur_y]), float(nfe)


def odeint_grid(func, y0, t, *args, step_size):
    """odeint_grid, lib/ode.py:561-577 / 750-754: ravel, _rk4_odeint, un-ravel with a leading time axis of 2."""
    flat0, unravel = ravel(y0)

    def fun(yflat, tt):
        return ravel(func(unravel(yflat), tt, *args))[0]

    ys, nfe = rk4_solve(fun, step_size, flat0, t)
    leaves_t = [unravel(ys[i]) for i in range(2)]
    l0, td = _flatten(leaves_t[0])
    stacked = [torch.stack([_flatten(lt)[0][j] for lt in leaves_t]) for j in range(len(l0))]
    return _unflatten(td, stacked), nfe


def odeint_grid_split_fwd(fwd_func, y0, t, *args, n_aux, step_size):
    """_rk4_odeint_aux / _sepaux_one / _sepaux (n_aux = 1 / 2 / 3), lib/ode.py:906-946: integrate y0[0] with fwd_func on
    the grid, zeros of shape [2, B, 1] for every auxiliary output."""
    out, nfe = odeint_grid(fwd_func, y0[0], t, *args, step_size=step_size)
    B = out.shape[1]
    return (out,) + tuple(torch.zeros((2, B, 1), dtype=out.dtype) for _ in range(n_aux)), nfe


def rk4_odeint_rev(func, step_size, ys, ts, args, g):
    """_rk4_odeint_rev, lib/ode.py:1531-1566, on a flat state: per segment t_bar = <func(ys[i]), g[i]>, then the same
    fixed-grid RK4 on (y, y_bar, t0_bar, args_bar) over [-ts[i], -ts[i-1]]."""
    def aug(state, tneg, *a):
        y, y_bar = state[0], state[1]
        y_dot, pull = torch.func.vjp(func, y, -tneg, *a)
        return (-y_dot, *pull(y_bar))

    y_bar = g[-1]
    t0_bar = torch.zeros((), dtype=ys.dtype)
    args_bar = tree_map(torch.zeros_like, tuple(args))
    rev_ts_bar = []
    for i in range(len(ts) - 1, 0, -1):
        t_bar = torch.dot(func(ys[i], ts[i], *args), g[i])
        t0_bar = t0_bar - t_bar
        out, _ = odeint_grid(aug, (ys[i], y_bar, t0_bar, args_bar), torch.stack([-ts[i], -ts[i - 1]]), *args, step_size=step_size)
        _, y_bar, t0_bar, args_bar = out
        y_bar, t0_bar, args_bar = tree_map(lambda x: x[1], (y_bar, t0_bar, args_bar))
        y_bar = y_bar + g[i - 1]
        rev_ts_bar.append(t_bar)
    ts_bar = torch.stack([t0_bar] + rev_ts_bar[::-1])
    return (y_bar, ts_bar, *args_bar)


def odeint_grid_split_rev(rev_func, step_size, res_ys, ts, args, g):
    """_rk4_odeint_sepaux_rev and siblings, lib/ode.py:1568-1593: flatten per time, rk4_odeint_rev with rev_func."""
    T = res_ys[0].shape[0]
    flat_res = torch.stack([ravel(tuple(r[i] for r in res_ys))[0] for i in range(T)])
    flat_g = torch.stack([ravel(tuple(x[i] for x in g))[0] for i in range(T)])
    _, unravel = ravel(tuple(r[0] for r in res_ys))

    def flat_func(yflat, tt, *a):
        return ravel(rev_func(unravel(yflat), tt, *a))[0]

    out = rk4_odeint_rev(flat_func, step_size, flat_res, ts, args, flat_g)
    return (unravel(out[0]), *out[1:])
