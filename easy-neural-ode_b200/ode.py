"""Host-side mirror of the reference's solver surface (lib/ode.py), over libeno.so.

Same names, argument order and return tuples as the reference:

    odeint(func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=inf) -> (ys, nfe)       lib/ode.py:540-559
    odeint_aux_one(fwd_func, rev_func, y0, t, *args, rtol, atol, mxstep)                 lib/ode.py:657-676
    odeint_sepaux(fwd_func, rev_func, y0, t, *args, rtol, atol, mxstep)                  lib/ode.py:678-697
    _heun_odeint / _bosh_odeint / ... (func, rtol, atol, mxstep, y0, ts, *args)          lib/ode.py:1048-1390

``func`` must be a dynamics-spec closure (dynamics.EnoFunc); an arbitrary Python
callable raises TypeError - the product path has no eager fallback.  The reference's
custom_vjp (lib/ode.py:1708, 1713) is a torch.autograd.Function whose backward is the
adjoint kernel path (eno_odeint_bwd).

Inputs may live on the host: they are copied to the current CUDA device for the call
and the results are returned on the input's device.
"""
import ctypes as C
import math
import os

import torch

from . import _cabi
from . import latent as _latent
from .dynamics import EnoFunc, GenDynamics

last_stats = {"fwd": None, "bwd": None}   # controller statistics of the most recent solves
_workspaces = {}


_dist = {"world": 1, "rank": 0}


def _nccl_path():
    try:
        import nvidia.nccl as _n
        import os
        p = os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2")
        return p.encode() if os.path.exists(p) else None
    except Exception:
        return None


def init_distributed(device=None):
    """Make the solves on this process' GPU collective over the default torch.distributed group
    (one process per GPU, batch sharded in equal contiguous row blocks, parameters replicated).
    After this call every rank must enter odeint / odeint_aux_one (and their backward) together; the
    ranks then follow the single-device step sequence of the concatenated batch and
    parameter / time gradients come back globally summed."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("init_distributed: call torch.distributed.init_process_group first")
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = _cabi.lib()
    h = _cabi.ctx(dev.index)
    path = _nccl_path()
    buf = (C.c_uint8 * 128)()
    if rank == 0 and world > 1:
        _cabi.check(L.eno_comm_unique_id(h, buf, path), h, "eno_comm_unique_id")
    box = [bytes(buf)]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    idb = (C.c_uint8 * 128).from_buffer_copy(box[0])
    with torch.cuda.device(dev):
        _cabi.check(L.eno_comm_init(h, idb, rank, world, path), h, "eno_comm_init")
    _dist["world"], _dist["rank"] = world, rank
    _dist["rows_global"] = None
    _dist["peer"] = False
    if world > 1 and world <= 8 and os.environ.get("ENO_EXCHANGE", "peer") != "nccl":
        # move the per-iteration exchange into the solver's finish kernels, over NVLink peer memory (include/eno.h):
        # every rank exports its exchange buffer as a CUDA IPC handle, the handles are gathered, every rank maps its peers
        nbytes = int(os.environ.get("ENO_PEER_BYTES", str(64 << 20)))
        mine = (C.c_uint8 * 64)()
        with torch.cuda.device(dev):
            rc = L.eno_peer_export(h, nbytes, mine)
        ok = [rc == 0]
        handles = [None] * world
        dist.all_gather_object(handles, (rc, bytes(mine)))
        if all(r == 0 for r, _ in handles):
            blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(b for _, b in handles))
            with torch.cuda.device(dev):
                rc = L.eno_peer_attach(h, blob)
            flags = [None] * world
            dist.all_gather_object(flags, rc)
            if all(f == 0 for f in flags):
                _dist["peer"] = True
            elif rc == 0:
                raise RuntimeError("init_distributed: a rank could not map its peers' exchange buffers; set ENO_EXCHANGE=nccl")
            else:
                _cabi.check(rc, h, "eno_peer_attach")
    _workspaces.clear()
    return world, rank


def shutdown_distributed(device=None):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _cabi.lib().eno_comm_destroy(_cabi.ctx(dev.index))
    _dist["world"], _dist["rank"] = 1, 0


def shard_rows(n_rows, world, rank):
    """Contiguous row blocks of the global batch, the first ``n_rows % world`` ranks one row longer (SURVEY.md section 8e:
    100 rows on 8 GPUs = 13, 13, 13, 13, 12, 12, 12, 12).  Uneven blocks need ``set_shards(n_rows)`` on every rank."""
    if n_rows < world:
        raise ValueError(f"global batch {n_rows} is smaller than the world size {world}")
    per, extra = divmod(n_rows, world)
    lo = rank * per + min(rank, extra)
    return slice(lo, lo + per + (1 if rank < extra else 0))


def set_shards(rows_global=None, device=None):
    """Declare the size of the concatenated batch of the collective solves (eno_comm_set_shards): ranks may then hold
    different numbers of rows (``shard_rows``).  ``None`` restores equal shards."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    h = _cabi.ctx(dev.index)
    world = _dist["world"]
    if rows_global is None:
        slot, rows = 0, 0
    else:
        slot, rows = -(-int(rows_global) // world), int(rows_global)
    _cabi.check(_cabi.lib().eno_comm_set_shards(h, slot, rows), h, "eno_comm_set_shards")
    _dist["rows_global"] = rows or None
    _workspaces.clear()


_deferred = {"on": False}


def set_deferred(on, fwd_iters=0, bwd_iters=0, device=None):
    """Switch the solves on this device between the default polled mode and the deferred mode of
    include/eno.h (fixed iteration budgets, no host synchronisation: CUDA-graph capturable)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    h = _cabi.ctx(dev.index)
    _cabi.check(_cabi.lib().eno_set_deferred(h, int(bool(on)), int(fwd_iters), int(bwd_iters)), h, "eno_set_deferred")
    _deferred["on"] = bool(on)


def _norm_device(device):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def workspace_tensor(kind, device=None):
    """The cached workspace ("fwd" or "bwd") of the most recent solves on the device (uint8 tensor whose
    first bytes are the device-side controller block)."""
    return _workspaces.get((_norm_device(device).index, kind))


def read_deferred_stats(kind, device=None):
    """Controller statistics of the last deferred solve; status -101 = iteration budget exhausted."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ws = workspace_tensor(kind, dev)
    st = _cabi.Stats()
    h = _cabi.ctx(dev.index)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(_cabi.lib().eno_read_stats(h, _ptr(ws), C.byref(st), stream), h, "eno_read_stats")
    return st.as_dict()


CTRL_DONE_I32 = 18      # int32 index of Ctrl.done / Ctrl.n_iters inside a workspace (csrc/eno_common.cuh)
CTRL_NITERS_I32 = 20


def _workspace(device, nbytes, kind="fwd"):
    key = (device.index, kind)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


class WorkspaceOwner:
    """Dedicated solve workspaces for a caller that keeps raw pointers into them alive - a captured CUDA graph
    (mnist_step.MnistNode).  The process-wide cache above hands its tensors to whoever asks next and replaces them
    when a later solve needs more bytes; an owner's workspaces are only ever touched through the owner, and ``gen``
    moves whenever one had to be re-allocated, so that the holder knows its captured pointers are stale."""

    def __init__(self, device):
        self.device, self.ws, self.gen = device, {}, 0

    def get(self, kind, nbytes):
        ws = self.ws.get(kind)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self.ws[kind] = ws
            self.gen += 1
        return ws


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _dev(t):
    if t.device.type == "cuda":
        return t.device
    if not torch.cuda.is_available():
        raise _cabi.EnoLibraryError("the ODE path needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _mx(mxstep):
    return 0 if (mxstep is None or math.isinf(mxstep)) else int(mxstep)


def _f32c(t, device):
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def _warn_status(what, status):
    """The reference has no error channel (lib/ode.py:829-831: the loop just ends); the library reports why."""
    if status > 0:
        import warnings
        why = {1: "mxstep reached", 2: "step size underflow (dt <= 0 or NaN)", 3: "non-finite error ratio"}.get(status, str(status))
        warnings.warn(f"{what}: the adaptive loop ended early ({why}); results are what the reference would return "
                      "(last interpolant extrapolated / NaNs)", RuntimeWarning, stacklevel=3)


def _grid_fwd_call(spec, y0, ts, pflat, step_size):
    """eno_odeint_grid_fwd (fixed-grid RK4, lib/ode.py:864-903) on device tensors. -> (ys [2,B,D], stats dict)."""
    L = _cabi.lib()
    device = y0.device
    h = _cabi.ctx(device.index)
    B, D = y0.shape
    T = ts.numel()
    desc = spec.desc(_cabi.AUG_NONE)
    nbytes = L.eno_workspace_bytes(C.byref(desc), B, T, 0, 0)
    ws = _workspace(device, nbytes)
    ys = torch.empty((T, B, D), dtype=torch.float32, device=device)
    st = _cabi.Stats()
    stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    with torch.cuda.device(device):
        rc = L.eno_odeint_grid_fwd(h, C.byref(desc), B, _ptr(y0), _ptr(ts), T, _ptr(pflat), None, float(step_size), _ptr(ws),
                                   ws.numel(), _ptr(ys), None, C.byref(st), stream)
    _cabi.check(rc, h, "eno_odeint_grid_fwd")
    return ys, st.as_dict()


def _fwd_call(spec, solver, y0, ts, pflat, rtol, atol, mxstep, trace_cap=0, owner=None):
    """eno_odeint_fwd on device tensors. -> (ys [T,B,D], stats dict, trace or None)."""
    L = _cabi.lib()
    device = y0.device
    h = _cabi.ctx(device.index)
    B, D = y0.shape
    T = ts.numel()
    desc = spec.desc(_cabi.AUG_NONE)
    nbytes = L.eno_workspace_bytes(C.byref(desc), B, T, _cabi.TABLEAUX[solver], 0)
    ws = owner.get("fwd", nbytes) if owner is not None else _workspace(device, nbytes)
    ys = torch.empty((T, B, D), dtype=torch.float32, device=device)
    trace = torch.zeros((trace_cap, 4), dtype=torch.float32, device=device) if trace_cap else None
    st = _cabi.Stats()
    stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    with torch.cuda.device(device):
        rc = L.eno_odeint_fwd(h, C.byref(desc), _cabi.TABLEAUX[solver], B, _ptr(y0), _ptr(ts), T, _ptr(pflat), None,
                              float(rtol), float(atol), _mx(mxstep), _ptr(ws), ws.numel(), _ptr(ys), None,
                              C.byref(st), _ptr(trace), trace_cap, stream)
    _cabi.check(rc, h, "eno_odeint_fwd")
    _warn_status("odeint (forward solve)", rc)
    return ys, st.as_dict(), trace


def _bwd_call(spec, aug, eps_in_args, ys, ts, pflat, g_y, g_r, rtol, atol, mxstep, trace_cap=0, eps=None, owner=None):
    """eno_odeint_bwd (aug NONE / REG) or eno_odeint_bwd_sepaux (aug FIN: g_r is [T, 2, B] and eps is read).
    -> (y0_bar, r0_bar [B] or [2, B], ts_bar, params_bar, stats, trace, eps_bar or None)."""
    L = _cabi.lib()
    device = ys.device
    h = _cabi.ctx(device.index)
    T, B, D = ys.shape
    fin = aug == _cabi.AUG_FIN
    desc = spec.desc(aug, eps_in_args, reg_order=spec.reg_order if aug == _cabi.AUG_REG else 0)
    nbytes = L.eno_workspace_bytes(C.byref(desc), B, T, 0, 1)
    ws = owner.get("bwd", nbytes) if owner is not None else _workspace(device, nbytes, "bwd")
    y0_bar = torch.empty((B, D), dtype=torch.float32, device=device)
    r0_bar = torch.empty((2, B) if fin else (B,), dtype=torch.float32, device=device)
    ts_bar = torch.empty((T,), dtype=torch.float32, device=device)
    p_bar = torch.empty((spec.n_params,), dtype=torch.float32, device=device)
    eps_bar = torch.empty((B, D), dtype=torch.float32, device=device) if fin else None
    trace = torch.zeros((trace_cap, 4), dtype=torch.float32, device=device) if trace_cap else None
    st = (_cabi.Stats * max(T - 1, 1))()
    stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    with torch.cuda.device(device):
        if fin:
            rc = L.eno_odeint_bwd_sepaux(h, C.byref(desc), B, _ptr(ys), _ptr(ts), T, _ptr(pflat), _ptr(eps), _ptr(g_y),
                                         _ptr(g_r), float(rtol), float(atol), _mx(mxstep), _ptr(ws), ws.numel(),
                                         _ptr(y0_bar), _ptr(r0_bar), _ptr(eps_bar), _ptr(ts_bar), _ptr(p_bar), st,
                                         _ptr(trace), trace_cap, stream)
        else:
            rc = L.eno_odeint_bwd(h, C.byref(desc), B, _ptr(ys), _ptr(ts), T, _ptr(pflat), _ptr(g_y), _ptr(g_r),
                                  float(rtol), float(atol), _mx(mxstep), _ptr(ws), ws.numel(), _ptr(y0_bar), _ptr(r0_bar),
                                  _ptr(ts_bar), _ptr(p_bar), st, _ptr(trace), trace_cap, stream)
    _cabi.check(rc, h, "eno_odeint_bwd")
    _warn_status("odeint (adjoint solve)", rc)
    return y0_bar, r0_bar, ts_bar, p_bar, [s.as_dict() for s in st][:max(T - 1, 0)], trace, eps_bar


class _OdeintFn(torch.autograd.Function):
    """custom_vjp pair (_odeint_fwd/_odeint_rev, lib/ode.py:1442-1444, 1494-1529) and the split
    variant (_odeint_aux_one_fwd/_rev, 1478-1480, 1677-1684) when ``aug`` is AUG_REG."""

    @staticmethod
    def forward(ctx, spec, aug, eps_in_args, rtol, atol, mxstep, y0, r0, ts, pflat):
        ys, st, _ = _fwd_call(spec, "dopri5", y0, ts, pflat, rtol, atol, mxstep)
        last_stats["fwd"] = st
        ctx.spec, ctx.aug, ctx.eps_in_args = spec, aug, eps_in_args
        ctx.tols = (rtol, atol, mxstep)
        ctx.save_for_backward(ys, ts, pflat)
        T, B, _ = ys.shape
        # aux outputs of the split wrappers are zeros of shape [T, B, 1] (lib/ode.py:998-1004)
        rs = torch.zeros((T, B, 1), dtype=ys.dtype, device=ys.device)
        if _deferred["on"]:   # not known yet (read_deferred_stats after the stream has drained)
            nfe = torch.full((), float("nan"), dtype=torch.float32, device=ys.device)
        else:
            nfe = torch.tensor(st["nfe"], dtype=torch.float32, device=ys.device)
        ctx.mark_non_differentiable(nfe)
        return ys, rs, nfe

    @staticmethod
    def backward(ctx, g_ys, g_rs, _g_nfe):
        ys, ts, pflat = ctx.saved_tensors
        rtol, atol, mxstep = ctx.tols
        T, B, D = ys.shape
        g_y = g_ys.contiguous().float()
        g_r = g_rs.reshape(T, B).contiguous().float()
        y0_bar, r0_bar, ts_bar, p_bar, st, _, _ = _bwd_call(ctx.spec, ctx.aug, ctx.eps_in_args, ys, ts, pflat, g_y, g_r,
                                                           rtol, atol, mxstep)
        last_stats["bwd"] = st
        return None, None, None, None, None, None, y0_bar, r0_bar, ts_bar, p_bar


class _OdeintSepauxFn(torch.autograd.Function):
    """custom_vjp pair of odeint_sepaux (_odeint_sepaux_fwd/_rev, lib/ode.py:1482-1484, 1686-1693) for the
    'fin' aug_dynamics: forward = the plain solve, two zero aux outputs; backward = the adjoint of
    (dy, dfro, dkin), which reads eps and carries eps_bar."""

    @staticmethod
    def forward(ctx, spec, rtol, atol, mxstep, y0, a0, b0, ts, eps, pflat):
        ys, st, _ = _fwd_call(spec, "dopri5", y0, ts, pflat, rtol, atol, mxstep)
        last_stats["fwd"] = st
        ctx.spec = spec
        ctx.tols = (rtol, atol, mxstep)
        ctx.save_for_backward(ys, ts, eps, pflat)
        T, B, _ = ys.shape
        z1 = torch.zeros((T, B, 1), dtype=ys.dtype, device=ys.device)   # lib/ode.py:1011-1018
        z2 = torch.zeros((T, B, 1), dtype=ys.dtype, device=ys.device)
        if _deferred["on"]:
            nfe = torch.full((), float("nan"), dtype=torch.float32, device=ys.device)
        else:
            nfe = torch.tensor(st["nfe"], dtype=torch.float32, device=ys.device)
        ctx.mark_non_differentiable(nfe)
        return ys, z1, z2, nfe

    @staticmethod
    def backward(ctx, g_ys, g_a, g_b, _g_nfe):
        ys, ts, eps, pflat = ctx.saved_tensors
        rtol, atol, mxstep = ctx.tols
        T, B, D = ys.shape
        g_y = g_ys.contiguous().float()
        g_aux = torch.stack([g_a.reshape(T, B), g_b.reshape(T, B)], dim=1).contiguous().float()   # [T, 2, B]
        y0_bar, r0_bar, ts_bar, p_bar, st, _, eps_bar = _bwd_call(ctx.spec, _cabi.AUG_FIN, True, ys, ts, pflat, g_y,
                                                                 g_aux, rtol, atol, mxstep, eps=eps)
        last_stats["bwd"] = st
        return None, None, None, None, y0_bar, r0_bar[0], r0_bar[1], ts_bar, eps_bar, p_bar


def _bwd_call_split(spec, aug, n_aux, ys, ts, pflat, eps, g_y, g_aux, rtol, atol, mxstep, grid_step=None):
    """eno_odeint_bwd_sepaux (or, with ``grid_step``, eno_odeint_grid_bwd) for a dynamics family whose rev_func carries
    ``n_aux`` auxiliary states (g_aux [T, n_aux, B]).  -> (y0_bar, aux0_bar [n_aux, B], ts_bar, params_bar, eps_bar, stats)."""
    L = _cabi.lib()
    device = ys.device
    h = _cabi.ctx(device.index)
    T, B, D = ys.shape
    desc = spec.desc(aug, True)
    nbytes = L.eno_workspace_bytes(C.byref(desc), B, T, 0, 1)
    ws = _workspace(device, nbytes, "bwd")
    y0_bar = torch.empty((B, D), dtype=torch.float32, device=device)
    a0_bar = torch.empty((n_aux, B), dtype=torch.float32, device=device)
    ts_bar = torch.empty((T,), dtype=torch.float32, device=device)
    p_bar = torch.empty((spec.n_params,), dtype=torch.float32, device=device)
    eps_bar = torch.empty((B, D), dtype=torch.float32, device=device)
    st = (_cabi.Stats * max(T - 1, 1))()
    stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    with torch.cuda.device(device):
        if grid_step is not None:
            rc = L.eno_odeint_grid_bwd(h, C.byref(desc), B, _ptr(ys), _ptr(ts), T, _ptr(pflat), _ptr(eps), _ptr(g_y), _ptr(g_aux),
                                       float(grid_step), _ptr(ws), ws.numel(), _ptr(y0_bar), _ptr(a0_bar), _ptr(eps_bar),
                                       _ptr(ts_bar), _ptr(p_bar), st, stream)
        else:
            rc = L.eno_odeint_bwd_sepaux(h, C.byref(desc), B, _ptr(ys), _ptr(ts), T, _ptr(pflat), _ptr(eps), _ptr(g_y),
                                         _ptr(g_aux), float(rtol), float(atol), _mx(mxstep), _ptr(ws), ws.numel(),
                                         _ptr(y0_bar), _ptr(a0_bar), _ptr(eps_bar), _ptr(ts_bar), _ptr(p_bar), st, None, 0, stream)
    _cabi.check(rc, h, "eno_odeint_grid_bwd" if grid_step is not None else "eno_odeint_bwd_sepaux")
    return y0_bar, a0_bar, ts_bar, p_bar, eps_bar, [s.as_dict() for s in st][:max(T - 1, 0)]


class _OdeintSplitFn(torch.autograd.Function):
    """custom_vjp pairs of odeint_sepaux / odeint_fin_sepaux for the FFJORD closures (lib/ode.py:1006-1032 primal,
    1482-1488 fwd, 1686-1693 rev; defvjp 1714-1715): forward = the plain solve of y0[0] with fwd_func, zeros for the
    n_aux auxiliary states; backward = the adjoint of rev_func on (y, aux...)."""

    @staticmethod
    def forward(ctx, spec, aug, n_aux, rtol, atol, mxstep, y0, ts, eps, pflat, *aux0):
        if rtol is None:      # fixed grid: atol carries step_size (odeint_grid_sepaux*, lib/ode.py:579-615)
            ys, st = _grid_fwd_call(spec, y0, ts, pflat, atol)
        else:
            ys, st, _ = _fwd_call(spec, "dopri5", y0, ts, pflat, rtol, atol, mxstep)
        last_stats["fwd"] = st
        ctx.spec, ctx.aug, ctx.n_aux = spec, aug, n_aux
        ctx.tols = (rtol, atol, mxstep)
        ctx.save_for_backward(ys, ts, eps, pflat)
        T, B, _ = ys.shape
        zs = tuple(torch.zeros((T, B, 1), dtype=ys.dtype, device=ys.device) for _ in range(n_aux))
        nfe = torch.tensor(st["nfe"], dtype=torch.float32, device=ys.device)
        ctx.mark_non_differentiable(nfe)
        return (ys,) + zs + (nfe,)

    @staticmethod
    def backward(ctx, g_ys, *rest):
        ys, ts, eps, pflat = ctx.saved_tensors
        rtol, atol, mxstep = ctx.tols
        T, B, D = ys.shape
        n_aux = ctx.n_aux
        g_y = g_ys.contiguous().float()
        g_aux = torch.stack([g.reshape(T, B) for g in rest[:n_aux]], dim=1).contiguous().float()   # [T, n_aux, B]
        y0_bar, a0_bar, ts_bar, p_bar, eps_bar, st = _bwd_call_split(ctx.spec, ctx.aug, n_aux, ys, ts, pflat, eps, g_y, g_aux,
                                                                      rtol, atol, mxstep, grid_step=atol if rtol is None else None)
        last_stats["bwd"] = st
        return (None,) * 6 + (y0_bar, ts_bar, eps_bar, p_bar) + tuple(a0_bar[q] for q in range(n_aux))


def _odeint_split(name, kind, fwd_func, rev_func, y0, t, args, rtol, atol, mxstep):
    spec = fwd_func.spec
    aug, n_aux = spec.AUG_OF_KIND[kind]
    eps, params = _split_args(rev_func, args)
    if eps is None:
        raise TypeError(f"{name}: the FFJORD closures read eps (ffjord_tabular.py:250-268)")
    if len(y0) != 1 + n_aux:
        raise TypeError(f"{name} integrates a state of {1 + n_aux} leaves with this rev_func; got {len(y0)}")
    y_in = y0[0]
    out_dev = y_in.device
    dev = _dev(y_in)
    y = _f32c(y_in.reshape(-1, spec.dim), dev)
    aux0 = tuple(_f32c(a.reshape(-1), dev) for a in y0[1:])
    ts = _f32c(torch.as_tensor(t), dev)
    epsd = _f32c(eps.reshape(-1, spec.dim), dev)
    pflat = _f32c(spec.flatten_params(params), dev)
    out = _OdeintSplitFn.apply(spec, aug, n_aux, rtol, atol, mxstep, y, ts, epsd, pflat, *aux0)
    if out_dev != out[0].device:
        out = tuple(o.to(out_dev) for o in out)
    return tuple(out[:-1]), out[-1]



def _split_args(func, args):
    if len(args) != len(func.arg_names):
        raise TypeError(f"{func!r} expects args {func.arg_names}, got {len(args)}")
    named = dict(zip(func.arg_names, args))
    return named.get("eps"), named["params"]


def _require_spec(func, what):
    if not isinstance(func, EnoFunc):
        raise TypeError(f"{what}: func must be a dynamics-spec closure (e.g. MLPDynamics(...).dynamics_wrap); "
                        "arbitrary Python callables cannot run inside the CUDA step kernel and there is no fallback")


def odeint(func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=math.inf):
    """Adaptive dopri5 solve; mirrors lib/ode.py:540-559.  -> (ys [T, B, D], nfe)."""
    _require_spec(func, "odeint")
    if func.kind != "plain":
        return _odeint_augmented(func, y0, t, args, rtol, atol, mxstep)
    spec = func.spec
    eps, params = _split_args(func, args)
    out_dev = y0.device
    dev = _dev(params if isinstance(params, torch.Tensor) else y0)
    y = _f32c(y0.reshape(-1, spec.dim), dev)
    ts = _f32c(torch.as_tensor(t), dev)
    pflat = _f32c(spec.flatten_params(params), dev)
    r0 = torch.zeros(y.shape[0], device=dev)
    ys, _, nfe = _OdeintFn.apply(spec, _cabi.AUG_NONE, eps is not None, rtol, atol, mxstep, y, r0, ts, pflat)
    ys = ys.reshape((ts.numel(),) + tuple(y0.shape))
    if out_dev != ys.device:
        ys, nfe = ys.to(out_dev), nfe.to(out_dev)
    return ys, nfe


def _odeint_augmented(func, y0, t, args, rtol, atol, mxstep, trace_cap=0):
    """odeint on aug_dynamics / all_aug_dynamics (evaluation path, mnist.py:333-334, 343-348): the state is
    the tuple (y, r) or (y, r, fro, kin); forward only (the scripts never differentiate this call).
    -> ((ys [T,B,D], r [T,B], ...), nfe), every auxiliary output with the shape of its y0 leaf."""
    spec = func.spec
    eps, params = _split_args(func, args)
    aug_const, n_aux = spec.AUG_OF_KIND[func.kind]
    if len(y0) != 1 + n_aux:
        raise TypeError(f"{func!r} integrates a state of {1 + n_aux} leaves, got {len(y0)}")
    y_in = y0[0]
    flat_engine = hasattr(spec, "eps_kinds") and spec.desc().arch in (_cabi.ARCH_CONCATSQUASH, _cabi.ARCH_CONCATCONV)
    if not flat_engine:
        for leaf in y0[1:]:
            if float(leaf.abs().max()) != 0.0:
                raise NotImplementedError("augmented solves start from zero auxiliary state (aug_init, mnist.py:416-438)")
    out_dev = y_in.device
    dev = _dev(y_in)
    y = _f32c(y_in.detach().reshape(-1, spec.dim), dev)
    ts = _f32c(torch.as_tensor(t), dev)
    pflat = _f32c(spec.flatten_params(params).detach(), dev)
    reads_eps = func.kind in spec.eps_kinds
    if reads_eps and eps is None:
        raise TypeError(f"{func!r} reads eps")
    epsd = _f32c(eps.reshape(-1, spec.dim), dev) if reads_eps else None
    L = _cabi.lib()
    h = _cabi.ctx(dev.index)
    B, D = y.shape
    T = ts.numel()
    desc = spec.desc(aug_const, eps is not None)
    nbytes = L.eno_workspace_bytes(C.byref(desc), B, T, 0, 0)
    ws = _workspace(dev, nbytes)
    ys = torch.empty((T, B, D), dtype=torch.float32, device=dev)
    aux = torch.zeros((T, n_aux, B), dtype=torch.float32, device=dev)
    if flat_engine:   # the flat-state engines start from the given auxiliary state (aux_out[0] is read, include/eno.h)
        for q in range(n_aux):
            aux[0, q] = _f32c(y0[1 + q].detach().reshape(-1), dev)
    trace = torch.zeros((trace_cap, 4), dtype=torch.float32, device=dev) if trace_cap else None
    st = _cabi.Stats()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        rc = L.eno_odeint_fwd(h, C.byref(desc), 0, B, _ptr(y), _ptr(ts), T, _ptr(pflat), _ptr(epsd), float(rtol),
                              float(atol), _mx(mxstep), _ptr(ws), ws.numel(), _ptr(ys), _ptr(aux), C.byref(st), _ptr(trace),
                              trace_cap, stream)
    _cabi.check(rc, h, "eno_odeint_fwd")
    last_stats["fwd"] = st.as_dict()
    if trace_cap:
        last_stats["fwd_trace"] = trace[: st.n_iters].cpu()
    outs = [ys.reshape((T,) + tuple(y_in.shape))] + [aux[:, q].reshape((T,) + tuple(y0[1 + q].shape)) for q in range(n_aux)]
    nfe = torch.tensor(st.nfe, dtype=torch.float32, device=out_dev)
    return tuple(o.to(out_dev) for o in outs), nfe


def odeint_aux_one(fwd_func, rev_func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=math.inf):
    """Split solve of lib/ode.py:657-676: forward integrates ``y0[0]`` with ``fwd_func`` and
    returns zeros for the auxiliary state; backward runs the adjoint with ``rev_func`` on the
    augmented state.  -> ((ys [T,B,D], rs [T,B,1]), nfe)."""
    _require_spec(fwd_func, "odeint_aux_one")
    _require_spec(rev_func, "odeint_aux_one")
    if fwd_func.kind != "plain" or rev_func.kind != "aug" or fwd_func.spec is not rev_func.spec:
        raise TypeError("odeint_aux_one expects (spec.fwd, spec.aug_dynamics) of the same dynamics spec")
    spec = fwd_func.spec
    eps, params = _split_args(rev_func, args)
    y_in, r_in = y0
    out_dev = y_in.device
    dev = _dev(y_in)
    y = _f32c(y_in.reshape(-1, spec.dim), dev)
    r0 = _f32c(r_in.reshape(-1), dev)
    ts = _f32c(torch.as_tensor(t), dev)
    pflat = _f32c(spec.flatten_params(params), dev)
    ys, rs, nfe = _OdeintFn.apply(spec, _cabi.AUG_REG, eps is not None, rtol, atol, mxstep, y, r0, ts, pflat)
    if out_dev != ys.device:
        ys, rs, nfe = ys.to(out_dev), rs.to(out_dev), nfe.to(out_dev)
    return (ys, rs), nfe


def odeint_sepaux(fwd_func, rev_func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=math.inf):
    """Split solve with two auxiliary states, lib/ode.py:678-697 (mnist.py:83, 326-332 with --reg_type fin):
    forward integrates ``y0[0]`` with ``fwd_func`` and returns zeros for both auxiliary states; backward
    runs the adjoint with ``rev_func`` = aug_dynamics('fin') on (y, fro, kin).
    -> ((ys [T,B,D], fro [T,B,1], kin [T,B,1]), nfe)."""
    _require_spec(fwd_func, "odeint_sepaux")
    _require_spec(rev_func, "odeint_sepaux")
    if getattr(fwd_func.spec, "num_layers", None) is not None:
        # FFJORD closures (ffjord_tabular.py:299-304): rev_func = aug_dynamics 'our' on (y, delta_logp, r)
        if fwd_func.kind != "plain" or rev_func.kind != "aug" or fwd_func.spec is not rev_func.spec:
            raise TypeError("odeint_sepaux expects (spec.fwd, spec.aug_dynamics) of one ConcatSquashDynamics built with reg_type='our'")
        return _odeint_split("odeint_sepaux", "aug", fwd_func, rev_func, y0, t, args, rtol, atol, mxstep)
    if fwd_func.kind != "plain" or rev_func.kind != "fin" or fwd_func.spec is not rev_func.spec:
        raise TypeError("odeint_sepaux expects (spec.fwd, spec.aug_dynamics) of one dynamics spec built with reg_type='fin'")
    spec = fwd_func.spec
    eps, params = _split_args(rev_func, args)
    if eps is None:
        raise TypeError("odeint_sepaux: the 'fin' dynamics read eps (mnist.py:294-300)")
    if len(y0) != 3:
        raise TypeError(f"odeint_sepaux integrates (y, fro, kin); got a state of {len(y0)} leaves")
    y_in, a_in, b_in = y0
    out_dev = y_in.device
    dev = _dev(y_in)
    y = _f32c(y_in.reshape(-1, spec.dim), dev)
    a0, b0 = _f32c(a_in.reshape(-1), dev), _f32c(b_in.reshape(-1), dev)
    ts = _f32c(torch.as_tensor(t), dev)
    epsd = _f32c(eps.reshape(-1, spec.dim), dev)
    pflat = _f32c(spec.flatten_params(params), dev)
    ys, fro, kin, nfe = _OdeintSepauxFn.apply(spec, rtol, atol, mxstep, y, a0, b0, ts, epsd, pflat)
    if out_dev != ys.device:
        ys, fro, kin, nfe = ys.to(out_dev), fro.to(out_dev), kin.to(out_dev), nfe.to(out_dev)
    return (ys, fro, kin), nfe


def odeint_fin_sepaux(fwd_func, rev_func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=math.inf):
    """Split solve with three auxiliary states, lib/ode.py:699-718 (jit wrapper 790-794, primal 1020-1032, fwd rule 1486-1488,
    rev rule shared with odeint_sepaux, defvjp 1715): what the FFJORD scripts call with ``--reg_type fin``
    (ffjord_mnist.py:78).  Forward integrates ``y0[0]`` with ``fwd_func`` and returns zeros for (delta_logp, fro, kin);
    backward runs the adjoint with ``rev_func`` = aug_dynamics('fin') = (dy, -div, mean (J eps)^2, mean f^2).
    -> ((ys [T,B,D], delta_logp [T,B,1], fro [T,B,1], kin [T,B,1]), nfe)."""
    _require_spec(fwd_func, "odeint_fin_sepaux")
    _require_spec(rev_func, "odeint_fin_sepaux")
    if (getattr(fwd_func.spec, "num_layers", None) is None or fwd_func.kind != "plain" or rev_func.kind != "fin"
            or fwd_func.spec is not rev_func.spec):
        raise TypeError("odeint_fin_sepaux expects (spec.fwd, spec.aug_dynamics) of one ConcatSquashDynamics built with reg_type='fin'")
    return _odeint_split("odeint_fin_sepaux", "fin", fwd_func, rev_func, y0, t, args, rtol, atol, mxstep)


def _require_flat(func, what):
    _require_spec(func, what)
    if getattr(func.spec, "num_layers", None) is None:
        raise NotImplementedError(f"{what}: the fixed-grid solver runs on the ConcatSquash / ConcatConv2D engine; MLPDynamics "
                                  "solves are adaptive (odeint, odeint_aux_one, odeint_sepaux)")


def odeint_grid(func, y0, t, *args, step_size):
    """Fixed-grid RK4, lib/ode.py:561-577 -> _rk4_odeint (864-903): steps of ``step_size`` from t[0], the last one clipped to
    land on t[-1]; returns the two rows (y0, y(t[-1])) and nfe = 4 per step.  Not differentiable through this wrapper
    (the scripts differentiate the split variants below).  -> (ys [2, B, D], nfe)."""
    _require_flat(func, "odeint_grid")
    if func.kind != "plain":
        raise NotImplementedError("odeint_grid: plain dynamics only (spec.fwd / spec.dynamics_wrap)")
    spec = func.spec
    _eps, params = _split_args(func, args)
    out_dev = y0.device
    dev = _dev(params if isinstance(params, torch.Tensor) else y0)
    y = _f32c(y0.reshape(-1, spec.dim), dev)
    ts = _f32c(torch.as_tensor(t), dev)
    pflat = _f32c(spec.flatten_params(params), dev)
    with torch.no_grad():
        ys, st = _grid_fwd_call(spec, y, ts, pflat, step_size)
    last_stats["fwd"] = st
    nfe = torch.tensor(st["nfe"], dtype=torch.float32, device=dev)
    ys = ys.reshape((2,) + tuple(y0.shape))
    return (ys.to(out_dev), nfe.to(out_dev)) if out_dev != ys.device else (ys, nfe)


def odeint_grid_sepaux_one(fwd_func, rev_func, y0, t, *args, step_size):
    """Fixed-grid split solve with TWO auxiliary states, lib/ode.py:598-615 -> _rk4_odeint_sepaux_one (906-918), adjoint
    _rk4_odeint_rev (1531-1566): what the FFJORD scripts use as ``_odeint_aux2`` with a fixed-grid method
    (ffjord_tabular.py:77, ffjord_mnist.py:70).  -> ((ys [2,B,D], delta_logp [2,B,1], r [2,B,1]), nfe)."""
    _require_flat(fwd_func, "odeint_grid_sepaux_one")
    _require_spec(rev_func, "odeint_grid_sepaux_one")
    if fwd_func.kind != "plain" or rev_func.kind != "aug" or fwd_func.spec is not rev_func.spec:
        raise TypeError("odeint_grid_sepaux_one expects (spec.fwd, spec.aug_dynamics) of one spec built with reg_type='our'")
    return _odeint_split("odeint_grid_sepaux_one", "aug", fwd_func, rev_func, y0, t, args, None, float(step_size), None)


def odeint_grid_sepaux(fwd_func, rev_func, y0, t, *args, step_size):
    """Fixed-grid split solve with THREE auxiliary states, lib/ode.py:579-596 -> _rk4_odeint_sepaux (934-946): ``_odeint_aux3``
    of the FFJORD scripts (--reg_type fin).  -> ((ys, delta_logp, fro, kin), nfe)."""
    _require_flat(fwd_func, "odeint_grid_sepaux")
    _require_spec(rev_func, "odeint_grid_sepaux")
    if fwd_func.kind != "plain" or rev_func.kind != "fin" or fwd_func.spec is not rev_func.spec:
        raise TypeError("odeint_grid_sepaux expects (spec.fwd, spec.aug_dynamics) of one spec built with reg_type='fin'")
    return _odeint_split("odeint_grid_sepaux", "fin", fwd_func, rev_func, y0, t, args, None, float(step_size), None)


def odeint_vmap(func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=math.inf):
    """``jax.vmap(lambda y0, t, *args: odeint(func, y0, t, *args), (0, None, None))`` - the per-example solves of
    mnist.py:350-353 (``--vmap`` NFE count): every row of ``y0`` is integrated as its own problem, with its own
    initial step, accept / reject sequence and NFE.  One kernel launch runs all of them to completion.
    Forward only (the script never differentiates it).  -> (ys [T=2, B, D], nfe [B]).

    For ``GenDynamics`` closures it is the per-trajectory solve of latent_ode.py:337-350,
    ``jax.vmap(lambda z_, t_: odeint(dynamics, init_fn(z_), t_, params), in_axes=(0, None))`` (any number of leading
    batch axes = nested vmaps): ``y0`` = z [..., D] or (z, r) for ``aug_dynamics``, ``t`` any number of output times,
    differentiable wrt z, t and params -> (zs [..., T, D], nfe [...]) or ((zs, rs [..., T]), nfe [...])."""
    _require_spec(func, "odeint_vmap")
    if isinstance(func.spec, GenDynamics):
        # latent_ode.py:337-350: every trajectory its own differentiable multi-output solve (latent.py)
        _, params = _split_args(func, args)
        return _latent.odeint_traj(func, y0, t, params, rtol, atol, mxstep)
    if func.kind != "plain":
        raise TypeError("odeint_vmap: only the plain dynamics are vmapped by the reference (mnist.py:351)")
    spec = func.spec
    _, params = _split_args(func, args)
    out_dev = y0.device
    dev = _dev(y0)
    y = _f32c(y0.detach().reshape(-1, spec.dim), dev)
    ts = _f32c(torch.as_tensor(t), dev)
    if ts.numel() != 2:
        raise NotImplementedError("odeint_vmap: t must be [t0, t1]")
    pflat = _f32c(spec.flatten_params(params).detach(), dev)
    L = _cabi.lib()
    h = _cabi.ctx(dev.index)
    B, D = y.shape
    ys = torch.empty((2, B, D), dtype=torch.float32, device=dev)
    nfe = torch.empty((B,), dtype=torch.float32, device=dev)
    iters = torch.empty((B, 2), dtype=torch.int32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    desc = spec.desc(_cabi.AUG_NONE)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        rc = L.eno_odeint_fwd_vmap(h, C.byref(desc), B, _ptr(y), _ptr(ts), 2, _ptr(pflat), float(rtol), float(atol),
                                   _mx(mxstep), _ptr(ys), _ptr(nfe), _ptr(iters), _ptr(status), stream)
    _cabi.check(rc, h, "eno_odeint_fwd_vmap")
    last_stats["vmap"] = {"iters": iters[:, 0].cpu(), "accepted": iters[:, 1].cpu(), "status": status.cpu()}
    return ys.reshape((2,) + tuple(y0.shape)).to(out_dev), nfe.to(out_dev)


def _tableau_odeint(solver):
    def run(func, rtol, atol, mxstep, y0, ts, *args):
        _require_spec(func, f"_{solver}_odeint")
        spec = func.spec
        _, params = _split_args(func, args)
        out_dev = y0.device
        dev = _dev(y0)
        y = _f32c(y0.reshape(-1, spec.dim), dev)
        tsd = _f32c(torch.as_tensor(ts), dev)
        pflat = _f32c(spec.flatten_params(params).detach(), dev)
        ys, st, _ = _fwd_call(spec, solver, y.detach(), tsd, pflat, rtol, atol, mxstep)
        last_stats["fwd"] = st
        ys = ys.reshape((tsd.numel(),) + tuple(y0.shape))
        nfe = torch.tensor(st["nfe"], dtype=torch.float32, device=out_dev)
        return ys.to(out_dev), nfe
    run.__name__ = f"_{solver}_odeint"
    run.__doc__ = (f"Forward-only adaptive solve with the {solver} tableau on a flat or [B, D] state; mirrors "
                   f"lib/ode.py's _{solver}_odeint(func, rtol, atol, mxstep, y0, ts, *args) -> (ys, nfe).")
    return run


_dopri5_odeint = _tableau_odeint("dopri5")
_heun_odeint = _tableau_odeint("heun")
_fehlberg_odeint = _tableau_odeint("fehlberg")
_bosh_odeint = _tableau_odeint("bosh")
_rk_fehlberg_odeint = _tableau_odeint("rk_fehlberg")
_cash_karp_odeint = _tableau_odeint("cash_karp")
_owrenzen_odeint = _tableau_odeint("owrenzen")
_owrenzen5_odeint = _tableau_odeint("owrenzen5")
_tanyam_odeint = _tableau_odeint("tanyam")


def solve_with_trace(spec, solver, y0, ts, params, rtol, atol, mxstep=math.inf, trace_cap=4096):
    """Forward solve that also returns the controller trace [(t, dt, ratio, accepted)]."""
    dev = _dev(y0)
    y = _f32c(y0.reshape(-1, spec.dim), dev)
    tsd = _f32c(torch.as_tensor(ts), dev)
    pflat = _f32c(spec.flatten_params(params).detach(), dev)
    ys, st, trace = _fwd_call(spec, solver, y, tsd, pflat, rtol, atol, mxstep, trace_cap)
    return ys, st, trace[: st["n_iters"]].cpu()


def rhs_closure(func, mode, y, t, *args, y_bar=None, aux_bar=None):
    """One evaluation of a dynamics-spec closure (eno_rhs) for the FFJORD closures: mode 0 -> dy; mode 1 -> (dy, aux
    [n_aux, B]) = what ``func`` returns; mode 2 -> (dy, aux, vjp_y, vjp_t, vjp_params, vjp_eps), the pull-back of
    (y_bar [B,D], aux_bar [n_aux,B]) through ``func`` at (y, t) - the adjoint's right-hand side, lib/ode.py:1498-1504."""
    _require_spec(func, "rhs_closure")
    spec = func.spec
    eps, params = _split_args(func, args)
    L = _cabi.lib()
    dev = _dev(y)
    h = _cabi.ctx(dev.index)
    yd = _f32c(y.reshape(-1, spec.dim), dev)
    B, D = yd.shape
    pflat = _f32c(spec.flatten_params(params).detach(), dev)
    if func.kind == "plain":
        aug, n_aux = _cabi.AUG_NONE, 0
    else:
        aug, n_aux = spec.AUG_OF_KIND[func.kind]
    desc = spec.desc(aug, eps is not None)
    nbytes = L.eno_workspace_bytes(C.byref(desc), B, 2, 0, 1 if mode == 2 else 0)
    ws = _workspace(dev, nbytes, "bwd" if mode == 2 else "fwd")
    dy = torch.empty_like(yd)
    aux = torch.empty((max(n_aux, 1), B), device=dev)
    epsd = _f32c(eps.reshape(-1, D), dev) if eps is not None else None
    zb, eb = torch.empty_like(yd), torch.empty_like(yd)
    tbar = torch.empty(1, device=dev)
    pbar = torch.empty(spec.n_params, device=dev)
    yb = _f32c(y_bar.reshape(-1, D), dev) if y_bar is not None else None
    ab = _f32c(aux_bar.reshape(n_aux, B), dev) if aux_bar is not None else None
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        rc = L.eno_rhs(h, C.byref(desc), mode, B, _ptr(yd), float(t), _ptr(pflat), _ptr(epsd), _ptr(yb), _ptr(ab), _ptr(ws),
                       ws.numel(), _ptr(dy), _ptr(aux), _ptr(zb), _ptr(tbar), _ptr(pbar), _ptr(eb), stream)
    _cabi.check(rc, h, "eno_rhs")
    if mode == 0:
        return dy
    if mode == 1:
        return dy, aux[:n_aux]
    return dy, aux[:n_aux], zb, tbar, pbar, eb


def rhs(spec, mode, y, t, params, y_bar=None, r_bar=None, eps=None):
    """One evaluation of the integrated function (eno_rhs): mode 0 f, 1 (f, r_K), 2 the VJP.
    With ``eps`` the 'fin' dynamics are evaluated instead: aux = (dfro, dkin) [2, B], r_bar is [2, B] and
    mode 2 also returns the cotangent of eps."""
    L = _cabi.lib()
    dev = _dev(y)
    h = _cabi.ctx(dev.index)
    yd = _f32c(y.reshape(-1, spec.dim), dev)
    B, D = yd.shape
    pflat = _f32c(spec.flatten_params(params).detach(), dev)
    fin = eps is not None and mode != 0
    desc = spec.desc((_cabi.AUG_FIN if fin else _cabi.AUG_REG) if mode else _cabi.AUG_NONE, True)
    nbytes = L.eno_workspace_bytes(C.byref(desc), B, 2, 0, 1)
    ws = _workspace(dev, nbytes, "bwd")
    dy = torch.empty_like(yd)
    r = torch.empty((2, B) if fin else (B,), device=dev)
    epsd = _f32c(eps.reshape(-1, D), dev) if fin else None
    eb = torch.empty_like(yd) if fin else None
    zb = torch.empty_like(yd)
    tbar = torch.empty(1, device=dev)
    pbar = torch.empty(spec.n_params, device=dev)
    yb = _f32c(y_bar.reshape(-1, D), dev) if y_bar is not None else None
    rb = _f32c(r_bar.reshape(-1), dev) if r_bar is not None else None
    if fin and mode != 2:   # the fin dynamics exist as the adjoint RHS only: evaluate it with zero cotangents
        yb, rb = torch.zeros_like(yd), torch.zeros(2 * B, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        rc = L.eno_rhs(h, C.byref(desc), 2 if fin else mode, B, _ptr(yd), float(t), _ptr(pflat), _ptr(epsd), _ptr(yb),
                       _ptr(rb), _ptr(ws), ws.numel(), _ptr(dy), _ptr(r), _ptr(zb), _ptr(tbar), _ptr(pbar), _ptr(eb),
                       stream)
    _cabi.check(rc, h, "eno_rhs")
    if mode == 0:
        return dy
    if mode == 1:
        return dy, r
    return (dy, r, zb, tbar, pbar, eb) if fin else (dy, r, zb, tbar, pbar)
