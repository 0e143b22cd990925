"""Thin driver reproducing one ``update()`` of mnist.py (553-559) around the ODE path:
PreODE (173-192) -> odeint_aux_one (330-341) -> PostODE (225-238) -> CE + lam*mean(r) (462-471)
-> jax.grad (528) -> momentum SGD (530-540).  With reg_type='fin' the ODE block is odeint_sepaux and the
loss is CE + lam_fro*mean(fro) + lam_kin*mean(kin) (mnist.py:83, 472-478).

Only what is needed to drive the hot path and to report the headline metric (train steps/s);
the head (sigmoid + Linear(10) on [B, 784]) is a handful of tiny torch ops, everything inside the
ODE block is libeno.so.
"""
import torch

from .dynamics import MLPDynamics
from . import ode


class MnistNode:
    def __init__(self, reg="r3", lam=6e-5, rtol=1.4e-8, atol=1.4e-8, device="cuda", dim=784, hidden=100,
                 lr=1e-1, mass=0.9, n_cls=10, world=1, reg_type="our", lam_fro=0.0, lam_kin=0.0):
        self.spec = MLPDynamics(dim, hidden, reg=reg, reg_type=reg_type)
        self.lam, self.rtol, self.atol = lam, rtol, atol
        self.reg_type, self.lam_fro, self.lam_kin = reg_type, lam_fro, lam_kin
        self.device = torch.device(device)
        self.lr, self.mass = lr, mass
        self.dim, self.n_cls = dim, n_cls
        self.world = world   # data-parallel ranks: the loss is the mean over the GLOBAL batch
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
        """Copy of the trainable state (params + momentum)."""
        return [t.clone() for t in (self.p_ode, self.post_w, self.post_b, self.v_ode, self.v_w, self.v_b)]

    def restore(self, snap):
        for dst, src in zip((self.p_ode, self.post_w, self.post_b, self.v_ode, self.v_w, self.v_b), snap):
            dst.copy_(src)

    def _reset_velocity(self):
        self.v_ode = torch.zeros_like(self.p_ode)
        self.v_w = torch.zeros_like(self.post_w)
        self.v_b = torch.zeros_like(self.post_b)

    def loss_and_grads(self, images_u8, labels, eps):
        """images [B,28,28,1] uint8, labels [B] int64, eps [B,784] Rademacher; all on self.device."""
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

    def update(self, images_u8, labels, eps):
        """One training step; returns the (device) loss tensor."""
        out = self.loss_and_grads(images_u8, labels, eps)
        g = out["grads"]
        self.apply_grads(g["ode"], g["post_w"], g["post_b"])
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
            self._g_in = tuple(t.clone() for t in (images_u8, labels, eps))
            try:
                lf, lb = self._level(self._last[0]), self._level(self._last[1])
                for key in ((lf, lb), (lf, 2 * lb), (2 * lf, 2 * lb), (2 * lf, 4 * lb), (4 * lf, 4 * lb)):
                    self._capture(key)
            except Exception as e:                           # capture is an optimisation, never a requirement
                self._graph_off, self._graph_error = True, repr(e)
            return dict(loss=float(out["loss"]), redone=False, fwd_iters=self._last[0], bwd_iters=self._last[1],
                        nfe=out["fwd_stats"]["nfe"])
        key = (self._level(self._last[0]), self._level(self._last[1]))
        if key not in self._graphs:
            self._capture(key)
        graph, info, self._launches = self._graphs[key]
        for dst, src in zip(self._g_in, (images_u8, labels, eps)):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        loss, ok_f, n_f, ok_b, n_b = info.tolist()           # the step's one device -> host read
        if not (ok_f and ok_b):
            # an iteration budget was too small: the guarded update left the parameters untouched; redo the
            # step in polled mode and move to the level it needs
            out = self.update(images_u8, labels, eps)
            self._last = (out["fwd_stats"]["n_iters"], out["bwd_stats"][0]["n_iters"])
            return dict(loss=float(out["loss"]), redone=True, fwd_iters=self._last[0], bwd_iters=self._last[1],
                        nfe=out["fwd_stats"]["nfe"])
        self._last = (int(n_f), int(n_b))
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
        finally:
            ode.set_deferred(False, device=self.device)
        self.restore(snap)

    def _graph_body(self):
        out = self.loss_and_grads(*self._g_in)
        wf = ode.workspace_tensor("fwd", self.device).view(torch.int32)
        wb = ode.workspace_tensor("bwd", self.device).view(torch.int32)
        ok = (wf[ode.CTRL_DONE_I32] != 0) & (wb[ode.CTRL_DONE_I32] != 0)
        g = out["grads"]
        for p, v, gr in ((self.p_ode, self.v_ode, g["ode"]), (self.post_w, self.v_w, g["post_w"]),
                         (self.post_b, self.v_b, g["post_b"])):
            v_new = self.mass * v + gr
            v.copy_(torch.where(ok, v_new, v))                # NaN-safe: garbage gradients are never selected
            p.copy_(torch.where(ok, p - self.lr * v_new, p))
        return torch.stack([out["loss"].float(), wf[ode.CTRL_DONE_I32].float(), wf[ode.CTRL_NITERS_I32].float(),
                            wb[ode.CTRL_DONE_I32].float(), wb[ode.CTRL_NITERS_I32].float()])

    def apply_grads(self, g_ode, g_w, g_b):
        """jax.experimental.optimizers.momentum: v <- mass*v + g ; x <- x - lr*v."""
        self.v_ode.mul_(self.mass).add_(g_ode); self.p_ode.sub_(self.lr * self.v_ode)
        self.v_w.mul_(self.mass).add_(g_w); self.post_w.sub_(self.lr * self.v_w)
        self.v_b.mul_(self.mass).add_(g_b); self.post_b.sub_(self.lr * self.v_b)
