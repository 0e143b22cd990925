"""Thin driver of the FORWARD (evaluation) path of ffjord_mnist.py around the ODE block: ForwardPreODE (the logit
pre-flow, 155-168, logdetgrad 138-144) -> all_ode / forward_all (326-338, 349-360, 409-417: odeint on all_aug_dynamics
from (y, logpy, 0, 0, 0)) -> bits per dim (455-470), and nfe_fn (362-385: odeint on ffjord_dynamics).

Training: one ``update()`` of the script (570-583): forward_aux (399-407: eps Rademacher, pre-flow, odeint_sepaux /
odeint_fin_sepaux from (y, logpy, 0[, 0])) -> loss_fn (480-497: bits per dim + lam mean(r) + lam_w 0.5 |theta|^2) ->
jax.grad -> clip_grads(max_grad_norm) -> Adam with the warm-up learning rate (548-556).  As in the reference the
training forward integrates y alone and the auxiliary OUTPUTS of the split solve are zeros (lib/ode.py:1006-1032), so
the loss value carries delta_logp = 0 and r = 0; their cotangents drive the adjoint, where the Hutchinson term and
the Taylor regulariser are evaluated.

The pre-flow, the loss head and the gradient clipping are element-wise torch ops on the device; the solves run on the
ConcatConv2D kernels (csrc/ccv_kernels.cuh), Adam is eno_adam_update.
"""
import ctypes as C
import math

import torch

from .dynamics import ConcatConvDynamics
from . import ode, _cabi


class ImageFfjord:
    def __init__(self, image=(28, 28), hidden=64, reg="r2", reg_type="our", rtol=1.4e-8, atol=1.4e-8, alpha=1e-5,
                 device="cuda", seed=0, lam=3e-4, lam_fro=0.0, lam_kin=0.0, lam_w=0.0, lr=1e-3, warmup_itrs=1e3,
                 max_grad_norm=1e10, world=1, rows_global=None):
        """world > 1: batch rows sharded over the ranks, ode.init_distributed() has run - the solves are ONE collective solve
        of the concatenated batch (css_engine.cu) and the parameter gradient they return is already the global one;
        rows_global: rows of the concatenated batch when the shards are uneven (ode.set_shards)."""
        self.spec = ConcatConvDynamics(image, hidden, reg=reg, reg_type=reg_type)
        self.image, self.rtol, self.atol, self.alpha = tuple(image), rtol, atol, alpha
        self.reg_type, self.lam, self.lam_fro, self.lam_kin, self.lam_w = reg_type, lam, lam_fro, lam_kin, lam_w
        self.lr, self.warmup_itrs, self.max_grad_norm, self.world = lr, warmup_itrs, max_grad_norm, world
        self.device = ode._norm_device(device)
        self.ts = torch.tensor([0., 1.], device=self.device)
        self.params = self.spec.flatten_params(self.spec.init_params(seed=seed, device=self.device)).clone()
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.itr = 0
        self.rows_global = rows_global

    def snapshot(self):
        return [t.clone() for t in (self.params, self.m, self.v)] + [self.itr]

    def restore(self, snap):
        for dst, src in zip((self.params, self.m, self.v), snap[:3]):
            dst.copy_(src)
        self.itr = snap[3]

    def loss_and_grads(self, images, eps):
        """images in [0, 1], eps Rademacher, both [B, H, W, 1] on self.device.  Forward solve -> loss head -> adjoint."""
        B = images.shape[0]
        n = self.rows_global or B * self.world
        if self.world > 1 and ode._dist["world"] != self.world:
            raise RuntimeError(f"world = {self.world} needs collective solves: call easy_neural_ode_b200.init_distributed() first "
                               "(otherwise every rank would return the gradient of its own rows only)")
        D = self.spec.dim
        fin = self.reg_type == "fin"
        aug, n_aux = self.spec.AUG_OF_KIND["fin" if fin else "aug"]
        y, _logpy = self.pre_ode(images, torch.zeros(B, 1, device=self.device))
        ys, st_f, _ = ode._fwd_call(self.spec, "dopri5", y.reshape(B, D), self.ts, self.params, self.rtol, self.atol, math.inf)
        z = ys[1]
        # _loss_fn (455-470) with delta_logp = 0 (the split solve's auxiliary outputs are zeros)
        logpz = (-0.5 * math.log(2 * math.pi) - z.square() / 2).sum() / (n * D)
        if self.world > 1:                                   # the mean runs over the rows of the concatenated batch
            import torch.distributed as dist
            dist.all_reduce(logpz)
        loss = -(logpz - math.log(256)) / math.log(2) + self.lam_w * 0.5 * self.params.square().sum()
        g_y = torch.zeros_like(ys)
        g_y[1] = z / (n * D * math.log(2))
        g_aux = torch.zeros((2, n_aux, B), device=self.device)
        g_aux[1, 0] = 1.0 / (n * D * math.log(2))           # d bits_per_dim / d delta_logp
        if fin:
            g_aux[1, 1], g_aux[1, 2] = self.lam_fro / n, self.lam_kin / n
        else:
            g_aux[1, 1] = self.lam / n
        _, _, ts_bar, p_bar, eps_bar, st_b = ode._bwd_call_split(self.spec, aug, n_aux, ys, self.ts, self.params, eps.reshape(B, D),
                                                                g_y, g_aux, self.rtol, self.atol, math.inf)
        ode.last_stats["fwd"], ode.last_stats["bwd"] = st_f, st_b
        return dict(loss=loss, z=z, nfe=st_f["nfe"], grads=p_bar, fwd_stats=st_f, bwd_stats=st_b)

    def lr_schedule(self, itr, num_batches=300):
        """548-556: linear warm-up, /10 after epoch 80."""
        frac = min((itr + 1.0) / max(self.warmup_itrs, 1.0), 1.0)
        return self.lr * frac if itr // num_batches < 80 else self.lr / 10

    def apply_grads(self, g):
        g = g + self.lam_w * self.params                     # the weight term of loss_fn is inside jax.grad
        norm = g.norm()                                      # optimizers.clip_grads(grad, max_grad_norm), 574-575
        g = torch.where(norm < self.max_grad_norm, g, g * (self.max_grad_norm / norm))
        L = _cabi.lib()
        h = _cabi.ctx(self.device.index)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            rc = L.eno_adam_update(h, ode._ptr(self.params), ode._ptr(self.m), ode._ptr(self.v), ode._ptr(g.contiguous()),
                                   self.params.numel(), float(self.lr_schedule(self.itr)), 0.9, 0.999, 1e-8, self.itr, 0.0, stream)
        _cabi.check(rc, h, "eno_adam_update")
        self.itr += 1

    def update(self, images, eps):
        out = self.loss_and_grads(images, eps)
        self.apply_grads(out["grads"])
        return out

    def pre_ode(self, x, logpx):
        """ForwardPreODE: y = logit(alpha + (1 - 2 alpha) x), logpy = logpx - logdetgrad(x, alpha)."""
        a = self.alpha
        s = a + (1 - 2 * a) * x
        logdet = (-torch.log(s - s * s) + math.log(1 - 2 * a)).reshape(x.shape[0], -1).sum(dim=1, keepdim=True)
        return torch.log(s) - torch.log(1 - s), logpx - logdet

    def forward_all(self, images, eps):
        """forward_all (409-417) + _loss_fn (455-470).  images in [0, 1], [B, H, W, 1]; eps Rademacher of the same shape."""
        B = images.shape[0]
        zero = torch.zeros(B, 1, device=self.device)
        y, logpy = self.pre_ode(images.to(self.device), zero)
        (ys, dl, rs, fro, kin), nfe = ode.odeint(self.spec.all_aug_dynamics, (y, logpy, zero, zero, zero), self.ts, eps, self.params,
                                                 rtol=self.rtol, atol=self.atol)
        z, delta_logp = ys[-1], dl[-1]
        logpz = (-0.5 * math.log(2 * math.pi) - z.square() / 2).reshape(B, -1).sum(dim=1, keepdim=True)
        logpx_per_dim = (logpz - delta_logp).sum() / z.numel()
        bpd = -(logpx_per_dim - math.log(256)) / math.log(2)
        return dict(z=z, delta_logp=delta_logp, r=rs[-1].reshape(-1), fro=fro[-1].reshape(-1), kin=kin[-1].reshape(-1),
                    nfe=float(nfe), bits_per_dim=float(bpd))

    def nfe(self, images, eps):
        """nfe_fn (362-385) without --vmap: one solve of ffjord_dynamics on (y, logpy) for the whole batch."""
        B = images.shape[0]
        zero = torch.zeros(B, 1, device=self.device)
        y, logpy = self.pre_ode(images.to(self.device), zero)
        return float(ode.odeint(self.spec.ffjord_dynamics, (y, logpy), self.ts, eps, self.params, rtol=self.rtol, atol=self.atol)[1])
