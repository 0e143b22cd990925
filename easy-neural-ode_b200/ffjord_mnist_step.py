"""Thin driver of the FORWARD (evaluation) path of ffjord_mnist.py around the ODE block: ForwardPreODE (the logit
pre-flow, 155-168, logdetgrad 138-144) -> all_ode / forward_all (326-338, 349-360, 409-417: odeint on all_aug_dynamics
from (y, logpy, 0, 0, 0)) -> bits per dim (455-470), and nfe_fn (362-385: odeint on ffjord_dynamics).

The pre-flow and the loss are element-wise torch ops on the device; the solves run on the ConcatConv2D kernels
(csrc/ccv_kernels.cuh).  Training (jax.grad through odeint_sepaux / odeint_fin_sepaux, 480-583) needs the adjoint of
the convolutional stack, which is not built.
"""
import math

import torch

from .dynamics import ConcatConvDynamics
from . import ode


class ImageFfjord:
    def __init__(self, image=(28, 28), hidden=64, reg="r2", reg_type="our", rtol=1.4e-8, atol=1.4e-8, alpha=1e-5,
                 device="cuda", seed=0):
        self.spec = ConcatConvDynamics(image, hidden, reg=reg, reg_type=reg_type)
        self.image, self.rtol, self.atol, self.alpha = tuple(image), rtol, atol, alpha
        self.device = ode._norm_device(device)
        self.ts = torch.tensor([0., 1.], device=self.device)
        self.params = self.spec.flatten_params(self.spec.init_params(seed=seed, device=self.device)).clone()

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
