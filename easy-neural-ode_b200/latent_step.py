"""Thin driver around the generative half of latent_ode.py's ``update()``: what happens between the sampled initial
latents ``z0 [3, 50, 20]`` and the loss (latent_ode.py:337-366, 441-457):

    z, r = vmap(vmap(odeint))(augment_dynamics(gen_dynamics), aug_init(z0), timesteps)      337-350   <- the ODE path
    pred = gen_to_data(z)                                                                    356   Linear(latent -> data)
    loss = -logsumexp_samples(likelihood(pred, data, mask)) + lam_gen * mean(r[..., -1])     404-409, 441-457
    grads -> Adam                                                                            480-500

The ODE-RNN encoder, rec_to_gen and the KL term (latent_ode.py:111-166, 312-336) are not part of
the path: the driver takes z0 as input, exactly what reaches the generative ODE.  Decoder, likelihood and Adam are a
handful of element-wise / [.., 20] x [20, 37] torch ops in float64 (host glue); the solves and their adjoints are the two
library launches of latent.py.  float64 like the script (jax_enable_x64, latent_ode.py:23).
"""
import math

import torch

from .dynamics import GenDynamics
from . import ode, latent


class LatentGen:
    def __init__(self, latent_dim=20, layers=3, units=50, data_dim=37, reg="r3", lam=1e-2, rtol=1.4e-8, atol=1.4e-8,
                 device="cuda", dtype=torch.float64, seed=0, lr=1e-2, scale=1.0, world=1):
        """world > 1 (torch.distributed initialised): the batch axis of z0 / data / mask is sharded over the ranks.  The
        per-trajectory solves are independent (vmap), so there is no collective inside the ODE path: the per-sample
        log-likelihood sums, the mask count and the regulariser mean are all-reduced (S + 2 scalars), every rank forms the
        same loss, and the parameter gradients are all-reduced once per step."""
        self.spec = GenDynamics(latent_dim, layers, units, reg=reg)
        self.device = ode._norm_device(device)
        self.dtype, self.lam, self.rtol, self.atol, self.lr = dtype, lam, rtol, atol, lr
        g = torch.Generator().manual_seed(seed + 100)
        self.params = self.spec.flatten_params(self.spec.init_params(seed=seed, device=self.device, dtype=dtype, scale=scale)).clone()
        # gen_to_data: hk.Linear(data_dim, w_init = RandomNormal(0.1)) (latent_ode.py:273-276, 304-307)
        self.dec_w = (0.1 * torch.randn((latent_dim, data_dim), generator=g, dtype=torch.float64)).to(self.device, dtype)
        self.dec_b = torch.zeros(data_dim, dtype=dtype, device=self.device)
        self.tensors = [self.params, self.dec_w, self.dec_b]
        self.m = [torch.zeros_like(t) for t in self.tensors]
        self.v = [torch.zeros_like(t) for t in self.tensors]
        self.itr = 0
        self.world = world

    def _global_sum(self, x):
        """sum over ranks as a value, identity as a derivative (every rank differentiates the same global loss wrt its own rows)"""
        if self.world == 1:
            return x
        import torch.distributed as dist
        tot = x.detach().clone()
        dist.all_reduce(tot)
        return x + (tot - x.detach())

    def snapshot(self):
        return [t.clone() for t in self.tensors + self.m + self.v] + [self.itr]

    def restore(self, snap):
        for dst, src in zip(self.tensors + self.m + self.v, snap[:-1]):
            dst.copy_(src)
        self.itr = snap[-1]

    def loss_and_grads(self, z0, ts, data, mask):
        """z0 [S, B, latent], ts [T], data / mask [B, T, data_dim] on self.device."""
        leaves = [t.detach().requires_grad_(True) for t in self.tensors]
        p, w, b = leaves
        r0 = torch.zeros(z0.shape[:-1], dtype=self.dtype, device=self.device)
        (zs, rs), nfe = ode.odeint_vmap(self.spec.aug_dynamics, (z0, r0), ts, p, rtol=self.rtol, atol=self.atol)
        pred = zs @ w + b                                               # [S, B, T, data_dim]
        std = 0.01
        ll = (-((data[None] * mask - pred * mask) ** 2) / (2 * std ** 2) - math.log(std) - math.log(2 * math.pi) / 2)
        ll = self._global_sum(ll.sum(dim=(1, 2, 3))) / self._global_sum(mask.sum())   # _likelihood, latent_ode.py:404-413
        n_traj = self._global_sum(torch.tensor(float(rs[..., -1].numel()), dtype=self.dtype, device=self.device))
        loss = -torch.logsumexp(ll, dim=0) + self.lam * self._global_sum(rs[..., -1].sum()) / n_traj
        grads = torch.autograd.grad(loss, leaves)
        if self.world > 1:
            import torch.distributed as dist
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat)
            out, o = [], 0
            for g in grads:
                out.append(flat[o:o + g.numel()].reshape(g.shape))
                o += g.numel()
            grads = tuple(out)
        return dict(loss=loss.detach(), nfe=nfe, grads=grads, fwd_iters=latent.last_stats["fwd"]["iters"],
                    bwd_iters=latent.last_stats["bwd"]["iters"])

    def apply_grads(self, grads, b1=0.9, b2=0.999, eps=1e-8):
        """jax.experimental.optimizers.adam (latent_ode.py:480-484), a few float64 torch ops."""
        i = self.itr
        for x, m, v, g in zip(self.tensors, self.m, self.v, grads):
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            mhat = m / (1 - b1 ** (i + 1))
            vhat = v / (1 - b2 ** (i + 1))
            x.sub_(self.lr * mhat / (vhat.sqrt() + eps))
        self.itr += 1

    def update(self, z0, ts, data, mask):
        out = self.loss_and_grads(z0, ts, data, mask)
        self.apply_grads(out["grads"])
        return out
