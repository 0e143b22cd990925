"""Thin driver reproducing one ``update()`` of mnist.py (553-559) around the ODE path:
PreODE (173-192) -> odeint_aux_one (330-341) -> PostODE (225-238) -> CE + lam*mean(r) (462-471)
-> jax.grad (528) -> momentum SGD (530-540).  With reg_type='fin' the ODE block is odeint_sepaux and the
loss is CE + lam_fro*mean(fro) + lam_kin*mean(kin) (mnist.py:83, 472-478).

Only what is needed to drive the hot path and to report the headline metric (train steps/s).  Everything
on the device is libeno.so: the ODE block, and - since the ~40 small framework launches of the head and the
optimiser were 8 % of a step - PostODE + loss + its gradients (eno_mnist_head) and the momentum update
(eno_momentum_update) too.  ``loss_and_grads_autograd`` keeps the torch.autograd formulation over the public
odeint_aux_one / odeint_sepaux surface; the tests hold the two against each other.
"""
import ctypes as C

import torch

from .dynamics import MLPDynamics
from . import ode, _cabi


def mnist_lr_schedule(num_batches=600):
    """lr_schedule of mnist.py:530-538: 1e-1 for epochs < 60, 1e-2 < 100, 1e-3 < 140, then 1e-4 (num_batches = 60000 /
    batch_size iterations per epoch)."""
    def lr(train_itr):
        epoch = train_itr // num_batches
        return 1e-1 if epoch < 60 else 1e-2 if epoch < 100 else 1e-3 if epoch < 140 else 1e-4
    return lr


class MnistNode:
    def __init__(self, reg="r3", lam=6e-5, rtol=1.4e-8, atol=1.4e-8, device="cuda", dim=784, hidden=100,
                 lr=1e-1, mass=0.9, n_cls=10, world=1, reg_type="our", lam_fro=0.0, lam_kin=0.0, lam_w=0.0):
        self.spec = MLPDynamics(dim, hidden, reg=reg, reg_type=reg_type)
        self.lam, self.rtol, self.atol = lam, rtol, atol
        self.reg_type, self.lam_fro, self.lam_kin = reg_type, lam_fro, lam_kin
        self.device = ode._norm_device(device)        # 'cuda' -> the current device, with an index
        # the solves of THIS model run in workspaces nobody else touches: the captured whole-step graphs hold raw
        # pointers into them (controller words, rings), so they must not be handed out or re-allocated behind them
        self._wsown = ode.WorkspaceOwner(self.device)
        self.lr, self.mass = lr, mass
        # lr: a number, or the script's schedule as a callable of the training iteration (mnist.py:530-538:
        # mnist_lr_schedule(num_batches)).  The kernel reads the step size from a device word, so the captured whole-step
        # graphs stay valid when the schedule moves.  lam_w: the weight-decay term of loss_fn (mnist.py:456-470).
        self.lam_w, self.itr = lam_w, 0
        self._lr_dev = torch.zeros(1, device=self.device)
        self._lr_val = None
        self.dim, self.n_cls = dim, n_cls
        self.world = world   # data-parallel ranks: the loss is the mean over the GLOBAL batch
        self.rows_global = None   # uneven shards (ode.set_shards): rows of the concatenated batch; None = world * B
        self.ts = torch.tensor([0., 1.], device=self.device)
        self.init_params(0)

    def init_params(self, seed=0, scale=1.0):
        p = self.spec.init_params(seed=seed, device=self.device, scale=scale)
        self.p_ode = self.spec.flatten_params(p).clone()
        g = torch.Generator().manual_seed(seed + 1)
        w = torch.empty((self.dim, self.n_cls), dtype=torch.float64)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
        self.post_w = (w / self.dim ** 0.5).float().to(self.device)
        self.post_b = torch.zeros(self.n_cls, device=self.device)
        self._reset_velocity()

    def load_params(self, p_ode, p_post):
        self.p_ode = self.spec.flatten_params({k: {kk: vv.to(self.device) for kk, vv in v.items()}
                                               for k, v in p_ode.items()}).clone()
        self.post_w = p_post["w"].to(self.device).clone()
        self.post_b = p_post["b"].to(self.device).clone()
        self._reset_velocity()

    def snapshot(self):
        """Copy of the trainable state (params + momentum + iteration counter)."""
        return [t.clone() for t in (self.p_ode, self.post_w, self.post_b, self.v_ode, self.v_w, self.v_b)] + [self.itr]

    def restore(self, snap):
        for dst, src in zip((self.p_ode, self.post_w, self.post_b, self.v_ode, self.v_w, self.v_b), snap):
            dst.copy_(src)
        self.itr = snap[6]

    def _reset_velocity(self):
        self.v_ode = torch.zeros_like(self.p_ode)
        self.v_w = torch.zeros_like(self.post_w)
        self.v_b = torch.zeros_like(self.post_b)

    # ---- library-only formulation (default) ---------------------------------------------------------
    def _buffers(self, B):
        """Per-batch-size device buffers of the fused head: cotangents handed to the adjoint, scratch, outputs."""
        if getattr(self, "_buf_B", None) == B:
            return self._buf
        D, Cn, dev = self.dim, self.n_cls, self.device
        fin = self.reg_type == "fin"
        g_y = torch.zeros((2, B, D), device=dev)                      # cotangent of ys: only the last time is hit
        n = self.rows_global or B * self.world
        if fin:
            g_r = torch.zeros((2, 2, B), device=dev)
            g_r[1, 0], g_r[1, 1] = self.lam_fro / n, self.lam_kin / n   # d(lam * mean(reg))/d reg[b], mnist.py:462-478
        else:
            g_r = torch.zeros((2, B), device=dev)
            g_r[1] = self.lam / n
        self._buf = dict(g_y=g_y, g_r=g_r, scratch=torch.empty(B * D + B * Cn + B, device=dev),
                         loss=torch.zeros(1, device=dev), gwb=torch.empty(D * Cn + Cn, device=dev))
        self._buf_B = B
        return self._buf

    def loss_and_grads(self, images_u8, labels, eps):
        """images [B,28,28,1] uint8, labels [B] int64, eps [B,784] Rademacher; all on self.device.
        forward solve -> eno_mnist_head -> adjoint solve; no autograd graph is built."""
        B = images_u8.shape[0]
        D, Cn = self.dim, self.n_cls
        buf = self._buffers(B)
        L = _cabi.lib()
        h = _cabi.ctx(self.device.index)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        x0 = (images_u8.to(torch.float32) / 255.).reshape(B, -1)
        ys, st_f, _ = ode._fwd_call(self.spec, "dopri5", x0, self.ts, self.p_ode, self.rtol, self.atol, float("inf"),
                                    owner=self._wsown)
        gw, gb = buf["gwb"][:D * Cn].view(D, Cn), buf["gwb"][D * Cn:]
        with torch.cuda.device(self.device):
            rc = L.eno_mnist_head(h, B, D, Cn, ode._ptr(ys[1]), ode._ptr(labels), ode._ptr(self.post_w), ode._ptr(self.post_b),
                                  1.0 / (self.rows_global or B * self.world), ode._ptr(buf["scratch"]), ode._ptr(buf["loss"]),
                                  ode._ptr(buf["g_y"][1]), ode._ptr(gw), ode._ptr(gb), stream)
        _cabi.check(rc, h, "eno_mnist_head")
        if self.world > 1:
            # the adjoint's params_bar is already the global sum (exchanged per iteration inside the
            # solve); the 7,850 head gradients are summed here
            import torch.distributed as dist
            dist.all_reduce(buf["gwb"])
        fin = self.reg_type == "fin"
        out = ode._bwd_call(self.spec, _cabi.AUG_FIN if fin else _cabi.AUG_REG, True, ys, self.ts, self.p_ode, buf["g_y"],
                            buf["g_r"], self.rtol, self.atol, float("inf"), eps=eps.contiguous() if fin else None,
                            owner=self._wsown)
        p_bar, st_b = out[3], out[4]
        ode.last_stats["fwd"], ode.last_stats["bwd"] = st_f, st_b
        if ode._deferred["on"]:
            nfe = torch.full((), float("nan"), device=self.device)
        else:
            nfe = torch.tensor(st_f["nfe"], dtype=torch.float32, device=self.device)
        return dict(loss=buf["loss"][0], y1=ys[1], nfe=nfe, grads={"ode": p_bar, "post_w": gw, "post_b": gb},
                    fwd_stats=st_f, bwd_stats=st_b)

    def _momentum(self, g_ode, g_w, g_b, ok_a=None, ok_b=None):
        """optimizers.momentum (mnist.py:530-540) for the three parameter blocks in one launch."""
        L = _cabi.lib()
        h = _cabi.ctx(self.device.index)
        vp = C.c_void_p
        ps = (vp * 3)(self.p_ode.data_ptr(), self.post_w.data_ptr(), self.post_b.data_ptr())
        vs = (vp * 3)(self.v_ode.data_ptr(), self.v_w.data_ptr(), self.v_b.data_ptr())
        gs = (vp * 3)(g_ode.data_ptr(), g_w.data_ptr(), g_b.data_ptr())
        ns = (C.c_int64 * 3)(self.p_ode.numel(), self.post_w.numel(), self.post_b.numel())
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            rc = L.eno_momentum_update(h, 3, ps, vs, gs, ns, 0.0, float(self.mass), float(self.lam_w),
                                       C.c_void_p(self._lr_dev.data_ptr()), ok_a, ok_b, stream)
        _cabi.check(rc, h, "eno_momentum_update")

    # ---- torch.autograd formulation over the public solver surface -----------------------------------
    def loss_and_grads_autograd(self, images_u8, labels, eps):
        """The same through odeint_aux_one / odeint_sepaux and torch.autograd (head and loss as torch ops)."""
        B = images_u8.shape[0]
        x0 = (images_u8.to(torch.float32) / 255.).reshape(B, -1)
        p_ode = self.p_ode.detach().requires_grad_(True)
        w = self.post_w.detach().requires_grad_(True)
        b = self.post_b.detach().requires_grad_(True)
        if self.reg_type == "fin":
            y0 = (x0, torch.zeros(B, device=self.device), torch.zeros(B, device=self.device))
            (ys, fro, kin), nfe = ode.odeint_sepaux(self.spec.fwd, self.spec.aug_dynamics, y0, self.ts, eps, p_ode,
                                                    rtol=self.rtol, atol=self.atol)
            y1 = ys[-1]
            reg_term = self.lam_fro * fro[-1].mean() + self.lam_kin * kin[-1].mean()
        else:
            y0 = (x0, torch.zeros(B, device=self.device))
            (ys, rs), nfe = ode.odeint_aux_one(self.spec.fwd, self.spec.aug_dynamics, y0, self.ts, eps, p_ode,
                                               rtol=self.rtol, atol=self.atol)
            y1 = ys[-1]
            reg_term = self.lam * rs[-1].mean()
        logits = torch.sigmoid(y1) @ w + b
        ce = torch.nn.functional.cross_entropy(logits, labels)
        # this rank's share of loss_fn (mnist.py:462-478) over the global batch of world * B samples
        loss = (ce + reg_term) / self.world
        g_ode, g_w, g_b = torch.autograd.grad(loss, (p_ode, w, b))
        if self.world > 1:
            # the adjoint's params_bar is already the global sum (exchanged per iteration inside the
            # solve); the tiny head gradients are summed here
            import torch.distributed as dist
            flat = torch.cat([g_w.reshape(-1), g_b])
            dist.all_reduce(flat)
            g_w, g_b = flat[:g_w.numel()].reshape(g_w.shape), flat[g_w.numel():]
        return dict(loss=ce.detach(), y1=y1.detach(), nfe=nfe, grads={"ode": g_ode, "post_w": g_w, "post_b": g_b},
                    fwd_stats=ode.last_stats["fwd"], bwd_stats=ode.last_stats["bwd"])

    def current_lr(self):
        return float(self.lr(self.itr)) if callable(self.lr) else float(self.lr)

    def _stage_lr(self):
        """Put the schedule's value for this iteration into the device word the update kernel reads."""
        lr = self.current_lr()
        if lr != self._lr_val:
            self._lr_dev.fill_(lr)
            self._lr_val = lr

    def update(self, images_u8, labels, eps):
        """One training step; returns the (device) loss tensor."""
        out = self.loss_and_grads(images_u8, labels, eps)
        g = out["grads"]
        self.apply_grads(g["ode"], g["post_w"], g["post_b"])      # stages the schedule's value at self.itr
        self.itr += 1
        return out

    # ---- whole-step CUDA graph ------------------------------------------------------------------
    # The solves normally poll their device-side controller once per chunk of iterations.  In deferred
    # mode (include/eno.h) they enqueue a fixed iteration budget and never synchronise, so the entire
    # update() - PreODE, forward solve, head, adjoint solve, momentum - becomes ONE graph launch.  The
    # parameter update is guarded on the device by the controllers' `done` flags; if a budget turns out too
    # small the parameters are left untouched and the step is redone in polled mode with a bigger budget.
    def update_graphed(self, images_u8, labels, eps):
        """update() through a captured CUDA graph; returns dict(loss=float, redone=bool, ...).

        Graphs are cached per pair of iteration-budget levels (powers of two); the level is chosen from the
        iteration counts of the previous step (they drift slowly).  The first call captures a small ladder
        of levels up front, the way a tracing compiler pays its compile time in the warm-up."""
        if getattr(self, "_graph_off", False):
            out = self.update(images_u8, labels, eps)
            return dict(loss=float(out["loss"]), redone=False, fwd_iters=out["fwd_stats"]["n_iters"],
                        bwd_iters=out["bwd_stats"][0]["n_iters"], nfe=out["fwd_stats"]["nfe"])
        if not hasattr(self, "_graphs"):
            out = self.update(images_u8, labels, eps)        # polled step: learns the iteration counts
            self._graphs, self._last = {}, (out["fwd_stats"]["n_iters"], out["bwd_stats"][0]["n_iters"])
            self._graph_gen = (self._wsown.gen, images_u8.shape[0])
            self._g_in = tuple(t.clone() for t in (images_u8, labels, eps))
            try:
                lf, lb = self._level(self._last[0]), self._level(self._last[1])
                for key in ((lf, lb), (lf, 2 * lb), (2 * lf, 2 * lb), (2 * lf, 4 * lb), (4 * lf, 4 * lb)):
                    self._capture(key)
            except Exception as e:                           # capture is an optimisation, never a requirement
                self._graph_off, self._graph_error = True, repr(e)
            return dict(loss=float(out["loss"]), redone=False, fwd_iters=self._last[0], bwd_iters=self._last[1],
                        nfe=out["fwd_stats"]["nfe"])
        if self._graph_gen != (self._wsown.gen, images_u8.shape[0]):
            # a workspace was re-allocated (or the batch size changed): every captured pointer is stale
            self._graphs = {}
            self._g_in = tuple(t.clone() for t in (images_u8, labels, eps))
        key = (self._level(self._last[0]), self._level(self._last[1]))
        if key not in self._graphs:
            self._capture(key)
        graph, info, self._launches = self._graphs[key]
        self._stage_lr()
        for dst, src in zip(self._g_in, (images_u8, labels, eps)):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        raw = info.cpu()                                     # the step's one device -> host read
        loss = float(raw[:1].view(torch.float32))
        span = ode.CTRL_NITERS_I32 - ode.CTRL_DONE_I32
        ok_f, n_f, ok_b, n_b = int(raw[1]), int(raw[1 + span]), int(raw[2 + span]), int(raw[2 + 2 * span])
        if not (ok_f and ok_b):
            # an iteration budget was too small: the guarded update left the parameters untouched; redo the
            # step in polled mode and move to the level it needs
            out = self.update(images_u8, labels, eps)        # (advances self.itr)
            self._last = (out["fwd_stats"]["n_iters"], out["bwd_stats"][0]["n_iters"])
            return dict(loss=float(out["loss"]), redone=True, fwd_iters=self._last[0], bwd_iters=self._last[1],
                        nfe=out["fwd_stats"]["nfe"])
        self._last = (int(n_f), int(n_b))
        self.itr += 1
        return dict(loss=loss, redone=False, fwd_iters=self._last[0], bwd_iters=self._last[1], nfe=2.0 + 6.0 * n_f)

    @staticmethod
    def _level(n):
        """Iteration budget for an observed count n: next power of two above n + 4 (at least 16).  Launches
        past termination cost ~3 us each; a level change costs a capture (milliseconds) once."""
        lvl = 16
        while lvl < n + 4:
            lvl *= 2
        return lvl

    def _capture(self, key):
        snap = self.snapshot()
        ode.set_deferred(True, key[0], key[1], self.device)
        try:
            if not self._graphs:
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._graph_body()                        # warm-up on a side stream (allocations)
                torch.cuda.current_stream(self.device).wait_stream(side)
                self.restore(snap)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                info = self._graph_body()
            launches = ode.last_stats["fwd"]["n_launches"] + sum(b["n_launches"] for b in ode.last_stats["bwd"])
            self._graphs[key] = (g, info, launches)
            self._graph_gen = (self._wsown.gen, self._g_in[0].shape[0])
        finally:
            ode.set_deferred(False, device=self.device)
        self.restore(snap)

    def _graph_body(self):
        out = self.loss_and_grads(*self._g_in)
        wf = self._wsown.ws["fwd"].view(torch.int32)
        wb = self._wsown.ws["bwd"].view(torch.int32)
        g = out["grads"]
        # guarded on the device by the two `done` words: garbage gradients never reach the parameters
        self._momentum(g["ode"], g["post_w"], g["post_b"],
                       C.c_void_p(wf.data_ptr() + 4 * ode.CTRL_DONE_I32), C.c_void_p(wb.data_ptr() + 4 * ode.CTRL_DONE_I32))
        # one launch: [loss bits | fwd done, -, n_iters | bwd done, -, n_iters] as int32 (decoded on the host)
        d, n = ode.CTRL_DONE_I32, ode.CTRL_NITERS_I32
        return torch.cat([out["loss"].reshape(1).view(torch.int32), wf[d:n + 1], wb[d:n + 1]])

    def apply_grads(self, g_ode, g_w, g_b):
        """jax.experimental.optimizers.momentum: v <- mass*v + g ; x <- x - lr*v."""
        self._stage_lr()
        self._momentum(g_ode.contiguous(), g_w.contiguous(), g_b.contiguous())
