"""Thin driver reproducing one ``update()`` of ffjord_tabular.py (547-553) around the ODE path:
forward_aux (349-355: eps ~ N(0,1), odeint_sepaux with the FFJORD closures, 299-312) -> loss_fn (417-432: negative mean
log-likelihood under the standard normal + lam * mean(r) + lam_w * 0.5 |theta|^2) -> jax.grad (521) -> Adam with the
piecewise-constant learning rate (525-529).

Only what is needed to drive the hot path and report train steps/s for BASELINE configs[1].  As in the reference the
training forward integrates y alone (the auxiliary outputs of odeint_sepaux are zeros, lib/ode.py:1006-1018), so the loss
VALUE carries delta_logp = 0 and r = 0; their cotangents (1/B and lam/B) drive the adjoint, where the Hutchinson term and
the Taylor regulariser are actually evaluated.  Loss head: three element-wise torch ops; optimiser: eno_adam_update.
"""
import ctypes as C
import math

import torch

from .dynamics import ConcatSquashDynamics
from . import ode, _cabi


def piecewise_constant(boundaries, values):
    """optimizers.piecewise_constant (ffjord_tabular.py:525-526)."""
    def lr(i):
        return values[sum(i >= b for b in boundaries)]
    return lr


class TabularFfjord:
    def __init__(self, dim=43, hdim_factor=20, num_layers=2, reg="r2", lam=1e-2, lam_w=1e-6, rtol=1.4e-8, atol=1.4e-8,
                 device="cuda", reg_type="our", lam_fro=0.0, lam_kin=0.0, world=1, seed=0,
                 lr_schedule=piecewise_constant([9000, 12750], [1e-3, 1e-4, 1e-5])):
        self.spec = ConcatSquashDynamics(dim, dim * hdim_factor, num_layers, reg=reg, reg_type=reg_type)
        self.dim, self.lam, self.lam_w, self.rtol, self.atol = dim, lam, lam_w, rtol, atol
        self.reg_type, self.lam_fro, self.lam_kin = reg_type, lam_fro, lam_kin
        self.device = ode._norm_device(device)
        # world > 1: rows sharded over the ranks and ode.init_distributed() has run - the solves are ONE collective solve of the
        # concatenated batch (css_engine.cu) and the parameter gradient they return is already the global one
        self.world = world
        self.lr_schedule = lr_schedule
        self.ts = torch.tensor([0., 1.], device=self.device)
        self.params = self.spec.flatten_params(self.spec.init_params(seed=seed, device=self.device)).clone()
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.itr = 0

    def snapshot(self):
        return [t.clone() for t in (self.params, self.m, self.v)] + [self.itr]

    def restore(self, snap):
        for dst, src in zip((self.params, self.m, self.v), snap[:3]):
            dst.copy_(src)
        self.itr = snap[3]

    def loss_and_grads(self, x, eps):
        """x, eps [B, dim] on self.device.  Forward solve -> loss head -> adjoint solve (no autograd graph)."""
        B = x.shape[0]
        n = B * self.world
        if self.world > 1 and ode._dist["world"] != self.world:
            raise RuntimeError(f"world = {self.world} needs collective solves: call easy_neural_ode_b200.init_distributed() first "
                               "(otherwise every rank would return the gradient of its own rows only)")
        fin = self.reg_type == "fin"
        aug, n_aux = self.spec.AUG_OF_KIND["fin" if fin else "aug"]
        ys, st_f, _ = ode._fwd_call(self.spec, "dopri5", x, self.ts, self.params, self.rtol, self.atol, math.inf)
        z = ys[1]
        # _loss_fn (402-407) with delta_logp = 0: -mean_b sum_d (-0.5 log 2 pi - z^2 / 2)
        zz = 0.5 * (z * z).sum() / n
        if self.world > 1:                                   # collective solves: the mean runs over the concatenated batch
            import torch.distributed as dist
            dist.all_reduce(zz)
        loss = 0.5 * self.dim * math.log(2 * math.pi) + zz
        g_y = torch.zeros_like(ys)
        g_y[1] = z / n
        g_aux = torch.zeros((2, n_aux, B), device=self.device)
        g_aux[1, 0] = 1.0 / n                               # d(-mean(logpz - delta_logp)) / d delta_logp
        if fin:
            g_aux[1, 1], g_aux[1, 2] = self.lam_fro / n, self.lam_kin / n
        else:
            g_aux[1, 1] = self.lam / n
        _, _, ts_bar, p_bar, eps_bar, st_b = ode._bwd_call_split(self.spec, aug, n_aux, ys, self.ts, self.params, eps, g_y, g_aux,
                                                                self.rtol, self.atol, math.inf)
        ode.last_stats["fwd"], ode.last_stats["bwd"] = st_f, st_b
        return dict(loss=loss, z=z, nfe=st_f["nfe"], grads=p_bar, fwd_stats=st_f, bwd_stats=st_b)

    def kernel_launches(self, out):
        """Library kernels enqueued by one update() (css_engine.cu: css_rhs_launch / enqueue_iteration / run_loop): an
        evaluation of the plain dynamics is L + 2 launches (stage set-up, L GEMMs, assemble), one of the adjoint's
        integrand 5 L + 6 (primal, tangent, two reverse and weight-gradient GEMMs per layer + seeds, column sums,
        assemble); an iteration is 6 evaluations + controller + dense output; each solve starts with 2 evaluations + 2
        norm kernels and a handful of set-up kernels; + Adam."""
        L = self.spec.num_layers + 1
        f_rhs, b_rhs = L + 2, 5 * L + 6
        n = 2 * (f_rhs + 1) + out["fwd_stats"]["n_launches"] * (6 * f_rhs + 2) + 2 * L + 4
        for st in out["bwd_stats"]:
            n += 2 * (b_rhs + 1) + st["n_launches"] * (6 * b_rhs + 2) + b_rhs + 6
        return n + 4 * L + 3 + 1

    def apply_grads(self, g):
        L = _cabi.lib()
        h = _cabi.ctx(self.device.index)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            rc = L.eno_adam_update(h, ode._ptr(self.params), ode._ptr(self.m), ode._ptr(self.v), ode._ptr(g), self.params.numel(),
                                   float(self.lr_schedule(self.itr)), 0.9, 0.999, 1e-8, self.itr, float(self.lam_w), stream)
        _cabi.check(rc, h, "eno_adam_update")
        self.itr += 1

    def update(self, x, eps):
        out = self.loss_and_grads(x, eps)
        self.apply_grads(out["grads"])
        return out
