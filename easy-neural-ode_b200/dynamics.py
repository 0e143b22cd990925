"""Dynamics-spec objects: what replaces the reference's Python ``func``.

A CUDA kernel cannot call a traced Python closure, so the callables the reference
scripts hand to ``odeint`` (mnist.py:285-334) become small descriptor objects that
``ode.odeint`` recognises and maps to a kernel family + flags.  Calling one directly
raises: there is deliberately no eager / CPU implementation on the product path.
"""
import math

import torch

from . import _cabi

L1 = "mlp_dynamics/linear"      # Haiku module names of MLPDynamics (mnist.py:205-206)
L2 = "mlp_dynamics/linear_1"
REGS = ["r2", "r3", "r4", "r5", "r6"]   # mnist.py:27


class EnoFunc:
    """One of the closures of init_model(): identified by kind and argument layout."""

    def __init__(self, spec, kind, arg_names):
        self.spec, self.kind, self.arg_names = spec, kind, tuple(arg_names)

    def __call__(self, *a, **k):
        raise RuntimeError(
            f"{self.spec.__class__.__name__}.{self.kind} exists only as a CUDA kernel; pass it to "
            "easy_neural_ode_b200.odeint / odeint_aux_one instead of calling it (no CPU fallback)")

    def __repr__(self):
        return f"<EnoFunc {self.kind}{self.arg_names} of {self.spec!r}>"


class MLPDynamics:
    """mnist.py:195-222: sigmoid -> [t, .] Linear(hidden) -> sigmoid -> [t, .] Linear(dim).

    ``reg`` / ``reg_type`` are the script flags ``--reg`` / ``--reg_type`` (mnist.py:39, 49).
    Closures (same names as in init_model):
      dynamics_wrap(x, t, params)                   mnist.py:285
      fwd(y, t, eps, params)                        the lambda at mnist.py:331
      aug_dynamics(yr, t, eps, params)              mnist.py:302-313  (both --reg_type arms)
      all_aug_dynamics(yr, t, eps, params)          mnist.py:315-324
    """

    def __init__(self, dim, hidden=100, reg="none", reg_type="our"):
        if dim % 4 or hidden % 4:
            raise ValueError("MLPDynamics kernels need dim % 4 == 0 and hidden % 4 == 0")
        if reg not in ["none"] + REGS:
            raise ValueError(f"--reg must be one of {['none'] + REGS}")
        if reg in REGS[2:]:
            raise NotImplementedError("Taylor orders above r3 are not built (BASELINE configs use r2/r3)")
        if reg_type not in ("our", "fin"):
            raise ValueError("--reg_type must be 'our' or 'fin' (mnist.py:49)")
        self.dim, self.hidden, self.reg, self.reg_type = dim, hidden, reg, reg_type
        self.reg_order = 0 if reg == "none" else REGS.index(reg) + 2
        self.dynamics_wrap = EnoFunc(self, "plain", ("params",))
        self.fwd = EnoFunc(self, "plain", ("eps", "params"))
        # reg_type 'our': (dy, r_K); 'fin': (dy, dfro, dkin) from a jvp with eps (mnist.py:302-313)
        self.aug_dynamics = EnoFunc(self, "aug" if reg_type == "our" else "fin", ("eps", "params"))
        self.all_aug_dynamics = EnoFunc(self, "all", ("eps", "params"))

    def __repr__(self):
        return f"MLPDynamics(dim={self.dim}, hidden={self.hidden}, reg={self.reg!r}, reg_type={self.reg_type!r})"

    @property
    def n_params(self):
        D, H = self.dim, self.hidden
        return (D + 1) * H + H + (H + 1) * D + D

    def desc(self, aug=_cabi.AUG_NONE, eps_in_args=False, reg_order=None):
        return _cabi.Desc(_cabi.ARCH_MLP_SIGMOID, self.dim, self.hidden,
                          self.reg_order if reg_order is None else reg_order, aug, int(eps_in_args))

    # ---- parameters: Haiku dict layout <-> flat vector in ravel_pytree order ---------------
    def init_params(self, seed=0, device="cpu", scale=1.0):
        """Haiku default init: truncated normal (+-2 sigma), sigma = 1/sqrt(fan_in); biases 0."""
        g = torch.Generator().manual_seed(seed)
        D, H = self.dim, self.hidden

        def tn(shape, fan_in):
            w = torch.empty(shape, dtype=torch.float64)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            return (w * (scale / math.sqrt(fan_in))).float().to(device)

        return {L1: {"w": tn((D + 1, H), D + 1), "b": torch.zeros(H, device=device)},
                L2: {"w": tn((H + 1, D), H + 1), "b": torch.zeros(D, device=device)}}

    def flatten_params(self, params):
        """dict -> [b1 | W1 | b2 | W2] (sorted-key ravel order); differentiable."""
        if isinstance(params, torch.Tensor):
            if params.numel() != self.n_params:
                raise ValueError(f"flat params must have {self.n_params} elements")
            return params.reshape(-1)
        return torch.cat([params[L1]["b"].reshape(-1), params[L1]["w"].reshape(-1),
                          params[L2]["b"].reshape(-1), params[L2]["w"].reshape(-1)])

    def unflatten_params(self, flat):
        D, H = self.dim, self.hidden
        o1, o2, o3 = H, H + (D + 1) * H, H + (D + 1) * H + D
        return {L1: {"b": flat[:o1], "w": flat[o1:o2].reshape(D + 1, H)},
                L2: {"b": flat[o2:o3], "w": flat[o3:].reshape(H + 1, D)}}
