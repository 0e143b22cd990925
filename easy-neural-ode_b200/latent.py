"""Per-trajectory differentiable solves of the generative ODE of latent_ode.py (BASELINE config C3), over libeno.so.

What the reference does (latent_ode.py:337-350): ``jax.vmap(lambda z_, t_: odeint(dynamics, init_fn(z_), t_, params))``
over the 50 series of a batch, itself under ``jax.vmap`` over the 3 latent samples - 150 independent odeint problems,
each with its own adaptive step sequence over the ~49 output times and, on the way back, its own adjoint solve
(lib/ode.py:1494-1529).  ``odeint_vmap`` (ode.py) dispatches here for ``GenDynamics`` closures: one kernel launch
runs every trajectory's forward solve, one more every trajectory's complete adjoint (eno_traj_fwd / eno_traj_bwd).

The element type follows the state tensor: float64 is the script's precision (jax_enable_x64, latent_ode.py:23).
There is no CPU / eager fallback.
"""
import ctypes as C

import torch

from . import _cabi

last_stats = {"fwd": None, "bwd": None}
_ws = {}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _dtype_code(dt):
    if dt == torch.float64:
        return _cabi.F64
    if dt == torch.float32:
        return _cabi.F32
    raise TypeError(f"per-trajectory solves run in float64 or float32, not {dt}")


def _workspace(dev, nbytes):
    w = _ws.get(dev)
    if w is None or w.numel() < nbytes:
        w = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _ws[dev] = w
    return w


def traj_fwd(spec, aug, z0, ts, pflat, rtol, atol, mxstep):
    """z0 [n, D] (device, contiguous) -> zs [n, T, D], rs [n, T] or None, nfe [n], iters [n, 2], status [n]."""
    dev, dt = z0.device, z0.dtype
    n, D = z0.shape
    T = ts.numel()
    zs = torch.empty((n, T, D), dtype=dt, device=dev)
    rs = torch.empty((n, T), dtype=dt, device=dev) if aug else None
    nfe = torch.empty((n,), dtype=dt, device=dev)
    iters = torch.empty((n, 2), dtype=torch.int32, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    desc = spec.desc(_cabi.AUG_REG if aug else _cabi.AUG_NONE)
    L, h = _cabi.lib(), _cabi.ctx(dev.index)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.eno_traj_fwd(h, C.byref(desc), _dtype_code(dt), n, _ptr(z0), _ptr(ts), T, _ptr(pflat), float(rtol), float(atol),
                            int(mxstep), _ptr(zs), _ptr(rs), _ptr(nfe), _ptr(iters), _ptr(status), stream)
    _cabi.check(rc, h, "eno_traj_fwd")
    return zs, rs, nfe, iters, status


def traj_bwd(spec, aug, zs, rs, ts, pflat, g_z, g_r, rtol, atol, mxstep, per_traj=False):
    dev, dt = zs.device, zs.dtype
    n, T, D = zs.shape
    P = spec.n_params
    desc = spec.desc(_cabi.AUG_REG if aug else _cabi.AUG_NONE)
    L, h = _cabi.lib(), _cabi.ctx(dev.index)
    code = _dtype_code(dt)
    nbytes = L.eno_traj_workspace_bytes(C.byref(desc), code, n)
    ws = _workspace(dev, nbytes)
    z0_bar = torch.empty((n, D), dtype=dt, device=dev)
    r0_bar = torch.empty((n,), dtype=dt, device=dev) if aug else None
    ts_bar = torch.empty((n, T), dtype=dt, device=dev)
    p_bar = torch.empty((P,), dtype=dt, device=dev)
    p_traj = torch.empty((n, P), dtype=dt, device=dev) if per_traj else None
    iters = torch.empty((n, max(T - 1, 1), 2), dtype=torch.int32, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.eno_traj_bwd(h, C.byref(desc), code, n, _ptr(zs), _ptr(rs), _ptr(ts), T, _ptr(pflat), _ptr(g_z), _ptr(g_r),
                            float(rtol), float(atol), int(mxstep), _ptr(ws), C.c_size_t(ws.numel()), _ptr(z0_bar), _ptr(r0_bar),
                            _ptr(ts_bar), _ptr(p_bar), _ptr(p_traj), _ptr(iters), _ptr(status), stream)
    _cabi.check(rc, h, "eno_traj_bwd")
    return z0_bar, r0_bar, ts_bar, p_bar, p_traj, iters, status


def traj_rhs(spec, z, pflat, z_bar=None, r_bar=None, reg_order=None):
    """Unit-test hook (eno_traj_rhs): f [n, D], mean(r_K^2) [n] and, with z_bar, (vjp wrt z [n, D], wrt params [n, P])."""
    dev, dt = z.device, z.dtype
    n, D = z.shape
    desc = spec.desc(_cabi.AUG_REG, reg_order=reg_order)
    f = torch.empty((n, D), dtype=dt, device=dev)
    rdot = torch.empty((n,), dtype=dt, device=dev)
    zbd = torch.empty((n, D), dtype=dt, device=dev) if z_bar is not None else None
    pb = torch.empty((n, spec.n_params), dtype=dt, device=dev) if z_bar is not None else None
    L, h = _cabi.lib(), _cabi.ctx(dev.index)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.eno_traj_rhs(h, C.byref(desc), _dtype_code(dt), n, _ptr(z), _ptr(pflat), _ptr(z_bar), _ptr(r_bar), _ptr(f),
                            _ptr(rdot), _ptr(zbd), _ptr(pb), stream)
    _cabi.check(rc, h, "eno_traj_rhs")
    return f, rdot, zbd, pb


class _TrajFn(torch.autograd.Function):
    """custom_vjp of the vmapped odeint (lib/ode.py:1442-1444, 1494-1529, 1708)."""

    @staticmethod
    def forward(ctx, spec, aug, rtol, atol, mxstep, z0, ts, pflat):
        zs, rs, nfe, iters, status = traj_fwd(spec, aug, z0, ts, pflat, rtol, atol, mxstep)
        ctx.spec, ctx.aug, ctx.cfg = spec, aug, (rtol, atol, mxstep)
        ctx.save_for_backward(zs, rs if aug else zs.new_empty(0), ts, pflat)
        last_stats["fwd"] = {"iters": iters, "status": status}
        ctx.mark_non_differentiable(nfe)
        if aug:
            return zs, rs, nfe
        return zs, nfe

    @staticmethod
    def backward(ctx, g_zs, *rest):
        zs, rs, ts, pflat = ctx.saved_tensors
        aug = ctx.aug
        rtol, atol, mxstep = ctx.cfg
        g_z = (g_zs if g_zs is not None else torch.zeros_like(zs)).contiguous()
        g_r = None
        if aug:
            g_r = (rest[0] if rest[0] is not None else torch.zeros_like(rs)).contiguous()
        z0_bar, r0_bar, ts_bar, p_bar, _, iters, status = traj_bwd(ctx.spec, aug, zs, rs if aug else None, ts, pflat, g_z, g_r,
                                                                   rtol, atol, mxstep)
        last_stats["bwd"] = {"iters": iters, "status": status}
        # ts is shared by the trajectories (in_axes=(0, None)): its cotangent is the sum over them
        return None, None, None, None, None, z0_bar, ts_bar.sum(0), p_bar


def odeint_traj(func, y0, t, params, rtol, atol, mxstep):
    """Implementation of ode.odeint_vmap for GenDynamics closures."""
    spec = func.spec
    aug = func.kind == "aug"
    if aug:
        if not (isinstance(y0, (tuple, list)) and len(y0) == 2):
            raise TypeError("odeint_vmap with aug_dynamics expects y0 = (z, r) (aug_init, latent_ode.py:352-355)")
        z, r0 = y0
        if bool((torch.as_tensor(r0) != 0).any()):
            raise NotImplementedError("the regulariser state starts at 0 (aug_init, latent_ode.py:352-355)")
    else:
        z = y0
    if z.shape[-1] != spec.latent_dim:
        raise ValueError(f"state width {z.shape[-1]} != latent_dim {spec.latent_dim}")
    out_dev = z.device
    if z.device.type != "cuda":
        if not torch.cuda.is_available():
            raise _cabi.EnoLibraryError("the ODE path needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = z.device
    dt = z.dtype
    _dtype_code(dt)
    lead = tuple(z.shape[:-1])
    z2 = z.to(dev).reshape(-1, spec.latent_dim).contiguous()
    ts = torch.as_tensor(t).to(device=dev, dtype=dt).contiguous()
    pflat = spec.flatten_params(params).to(device=dev, dtype=dt).contiguous()
    mx = 0 if (mxstep is None or mxstep == float("inf")) else int(mxstep)
    T = ts.numel()
    if aug:
        zs, rs, nfe = _TrajFn.apply(spec, True, float(rtol), float(atol), mx, z2, ts, pflat)
        return (zs.reshape(lead + (T, spec.latent_dim)).to(out_dev), rs.reshape(lead + (T,)).to(out_dev)), nfe.reshape(lead).to(out_dev)
    zs, nfe = _TrajFn.apply(spec, False, float(rtol), float(atol), mx, z2, ts, pflat)
    return zs.reshape(lead + (T, spec.latent_dim)).to(out_dev), nfe.reshape(lead).to(out_dev)
