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
                          self.reg_order if reg_order is None else reg_order, aug, int(eps_in_args), 0)

    # closure kind -> (ENO_AUG_*, number of auxiliary state leaves)
    AUG_OF_KIND = {"aug": (_cabi.AUG_REG, 1), "all": (_cabi.AUG_ALL, 3), "fin": (_cabi.AUG_FIN, 2)}
    eps_kinds = ("all", "fin")     # closures that read eps

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


class ConcatSquashDynamics:
    """NN_Dynamics of ffjord_tabular.py:163-193: ``num_layers`` hidden ConcatSquashLinear layers of width ``hidden`` and
    one back to ``dim``, softplus (``_logaddexp``, 92-103) between them:  u = (x W + b) * sigmoid(w_g t + b_g) + w_b t.

    ``reg`` / ``reg_type`` are the script flags (ffjord_tabular.py:38, 52); any ``reg != 'none'`` means the fixed K = 2
    Taylor term of the script's sol_recursive (116-136).  Closures, named as in init_model (235-297):
      dynamics_wrap(x, t, params)                235
      fwd(y, t, eps, params)                     the lambda at 303
      ffjord_dynamics((y, p), t, eps, params)    250-258   (dy, -div)
      aug_dynamics(ypr, t, eps, params)          270-284   (dy, -div, r) or, reg_type 'fin', (dy, -div, dfro, dkin)
      all_aug_dynamics(ypr, t, eps, params)      286-297   (dy, -div, r, dfro, dkin)
    Parameters: the Haiku dict of the module ({layer: {"w", "b", "w_gate", "b_gate", "w_bias"}}, the flattened form the
    oracle uses) or one flat vector in ravel order, per layer [b | W | w_bias | b_gate | w_gate] (sub-modules linear,
    linear_1 (hyper-bias, no bias term), linear_2 (hyper-gate) in sorted-key order).
    """

    def __init__(self, dim, hidden, num_layers=2, reg="none", reg_type="our"):
        if reg not in ["none"] + REGS:
            raise ValueError(f"--reg must be one of {['none'] + REGS}")
        if reg_type not in ("our", "fin"):
            raise ValueError("--reg_type must be 'our' or 'fin' (ffjord_tabular.py:52)")
        if not 1 <= num_layers <= 5:
            raise ValueError("ConcatSquashDynamics supports 1..5 hidden layers")
        self.dim, self.hidden, self.num_layers, self.reg, self.reg_type = dim, hidden, num_layers, reg, reg_type
        self.reg_order = 0 if reg == "none" else 2
        self.dims = [dim] + [hidden] * num_layers + [dim]
        self.dynamics_wrap = EnoFunc(self, "plain", ("params",))
        self.fwd = EnoFunc(self, "plain", ("eps", "params"))
        self.ffjord_dynamics = EnoFunc(self, "ffjord", ("eps", "params"))
        self.aug_dynamics = EnoFunc(self, "aug" if reg_type == "our" else "fin", ("eps", "params"))
        self.all_aug_dynamics = EnoFunc(self, "all", ("eps", "params"))

    AUG_OF_KIND = {"ffjord": (_cabi.AUG_FFJORD, 1), "aug": (_cabi.AUG_REG, 2), "fin": (_cabi.AUG_FIN, 3), "all": (_cabi.AUG_ALL, 4)}
    eps_kinds = ("ffjord", "aug", "fin", "all")

    def __repr__(self):
        return (f"ConcatSquashDynamics(dim={self.dim}, hidden={self.hidden}, num_layers={self.num_layers}, "
                f"reg={self.reg!r}, reg_type={self.reg_type!r})")

    @property
    def layer_names(self):
        return ["nn__dynamics/concat_squash_linear" + ("" if i == 0 else f"_{i}") for i in range(len(self.dims) - 1)]

    @property
    def n_params(self):
        return sum(i * o + 4 * o for i, o in zip(self.dims[:-1], self.dims[1:]))

    def desc(self, aug=_cabi.AUG_NONE, eps_in_args=False, reg_order=None):
        return _cabi.Desc(_cabi.ARCH_CONCATSQUASH, self.dim, self.hidden,
                          self.reg_order if reg_order is None else reg_order, aug, int(eps_in_args), self.num_layers)

    def init_params(self, seed=0, device="cpu", scale=1.0):
        """Haiku default init: truncated normal (+-2 sigma), sigma = 1/sqrt(fan_in) (fan_in = 1 for the two
        hyper-networks of the scalar t); biases 0."""
        g = torch.Generator().manual_seed(seed)

        def tn(shape, fan_in):
            w = torch.empty(shape, dtype=torch.float64)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            return (w * (scale / math.sqrt(fan_in))).float().to(device)

        params = {}
        for name, d_in, d_out in zip(self.layer_names, self.dims[:-1], self.dims[1:]):
            params[name] = {"w": tn((d_in, d_out), d_in), "b": torch.zeros(d_out, device=device),
                            "w_gate": tn((1, d_out), 1), "b_gate": torch.zeros(d_out, device=device),
                            "w_bias": tn((1, d_out), 1)}
        return params

    def flatten_params(self, params):
        if isinstance(params, torch.Tensor):
            if params.numel() != self.n_params:
                raise ValueError(f"flat params must have {self.n_params} elements")
            return params.reshape(-1)
        parts = []
        for n in self.layer_names:
            p = params[n]
            parts += [p["b"].reshape(-1), p["w"].reshape(-1), p["w_bias"].reshape(-1), p["b_gate"].reshape(-1),
                      p["w_gate"].reshape(-1)]
        return torch.cat(parts)

    def unflatten_params(self, flat):
        out, o = {}, 0
        for n, d_in, d_out in zip(self.layer_names, self.dims[:-1], self.dims[1:]):
            b = flat[o:o + d_out]; o += d_out
            w = flat[o:o + d_in * d_out].reshape(d_in, d_out); o += d_in * d_out
            wb = flat[o:o + d_out].reshape(1, d_out); o += d_out
            bg = flat[o:o + d_out]; o += d_out
            wg = flat[o:o + d_out].reshape(1, d_out); o += d_out
            out[n] = {"w": w, "b": b, "w_gate": wg, "b_gate": bg, "w_bias": wb}
        return out


class ConcatConvDynamics:
    """NN_Dynamics of ffjord_mnist.py:186-232: ConcatConv2D layers (123-135) 1 -> hidden -> hidden -> hidden -> 1 channels,
    3x3, stride 1, zero padding 1, the time concatenated as input channel 0 of every layer, softplus between the layers,
    on an ``image = (H, W)`` one-channel state (NHWC as in the script; any [B, H, W, 1] / [B, H*W] tensor).

    Closures, named as in init_model (268-338): ``dynamics_wrap``, ``fwd``, ``ffjord_dynamics`` ((dy, -div), 291-299, the
    ``nfe_fn`` solve 362-385), ``aug_dynamics`` ((dy, -div, r) or, reg_type 'fin', (dy, -div, dfro, dkin), 311-324: the
    rev_func of odeint_sepaux / odeint_fin_sepaux) and ``all_aug_dynamics`` (326-338, the ``forward_all`` solve).
    Parameters: the Haiku dict {layer: {"w" [3, 3, c_in + 1, c_out], "b" [c_out]}} or one flat vector
    in ravel order, per layer [b | w].
    """

    def __init__(self, image=(28, 28), hidden=64, reg="none", reg_type="our"):
        if reg not in ["none"] + REGS:
            raise ValueError(f"--reg must be one of {['none'] + REGS}")
        if reg_type not in ("our", "fin"):
            raise ValueError("--reg_type must be 'our' or 'fin' (ffjord_mnist.py:45)")
        if hidden not in (32, 64):
            raise ValueError("ConcatConvDynamics supports 32 or 64 hidden channels (the script uses 64)")
        self.image, self.hidden, self.reg, self.reg_type = tuple(image), hidden, reg, reg_type
        self.num_layers = 3          # hidden ConcatConv2D layers (marks the FFJORD closure family for the split solves)
        self.dim = image[0] * image[1]
        if self.dim % 4:
            raise ValueError("image height x width must be a multiple of 4")
        self.reg_order = 0 if reg == "none" else 2
        self.chans = [1, hidden, hidden, hidden, 1]
        self.dynamics_wrap = EnoFunc(self, "plain", ("params",))
        self.fwd = EnoFunc(self, "plain", ("eps", "params"))
        self.ffjord_dynamics = EnoFunc(self, "ffjord", ("eps", "params"))
        self.aug_dynamics = EnoFunc(self, "aug" if reg_type == "our" else "fin", ("eps", "params"))
        self.all_aug_dynamics = EnoFunc(self, "all", ("eps", "params"))

    AUG_OF_KIND = ConcatSquashDynamics.AUG_OF_KIND
    eps_kinds = ConcatSquashDynamics.eps_kinds

    def __repr__(self):
        return f"ConcatConvDynamics(image={self.image}, hidden={self.hidden}, reg={self.reg!r}, reg_type={self.reg_type!r})"

    @property
    def layer_names(self):
        return ["nn__dynamics/concat_conv2_d" + ("" if i == 0 else f"_{i}") + "/conv2_d" for i in range(4)]

    @property
    def n_params(self):
        return sum(o + 9 * (i + 1) * o for i, o in zip(self.chans[:-1], self.chans[1:]))

    def desc(self, aug=_cabi.AUG_NONE, eps_in_args=False, reg_order=None):
        return _cabi.Desc(_cabi.ARCH_CONCATCONV, self.dim, self.hidden, self.reg_order if reg_order is None else reg_order,
                          aug, int(eps_in_args), 3, self.image[0], self.image[1])

    def init_params(self, seed=0, device="cpu", scale=1.0, last_scale=0.0):
        """hk.Conv2D defaults (truncated normal, sigma = 1/sqrt(9 (c_in + 1)), zero biases); the script zero-initialises
        the last layer's w (ffjord_mnist.py:214-215) - ``last_scale`` = 0 reproduces that."""
        g = torch.Generator().manual_seed(seed)
        params = {}
        for i, (name, c_in, c_out) in enumerate(zip(self.layer_names, self.chans[:-1], self.chans[1:])):
            w = torch.empty((3, 3, c_in + 1, c_out), dtype=torch.float64)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            s = last_scale if i == 3 else scale
            params[name] = {"b": torch.zeros(c_out, device=device),
                            "w": (w * (s / math.sqrt(9 * (c_in + 1)))).float().to(device)}
        return params

    def flatten_params(self, params):
        if isinstance(params, torch.Tensor):
            if params.numel() != self.n_params:
                raise ValueError(f"flat params must have {self.n_params} elements")
            return params.reshape(-1)
        return torch.cat([t.reshape(-1) for n in self.layer_names for t in (params[n]["b"], params[n]["w"])])

    def unflatten_params(self, flat):
        out, o = {}, 0
        for n, c_in, c_out in zip(self.layer_names, self.chans[:-1], self.chans[1:]):
            b = flat[o:o + c_out]; o += c_out
            w = flat[o:o + 9 * (c_in + 1) * c_out].reshape(3, 3, c_in + 1, c_out); o += 9 * (c_in + 1) * c_out
            out[n] = {"b": b, "w": w}
        return out


class GenDynamics:
    """GenDynamics of latent_ode.py:191-208: ``[tanh, Linear(units)] x (layers + 1)`` then ``[tanh, Linear(latent_dim)]``,
    time-independent.  ``reg`` is the script flag ``--reg`` (latent_ode.py:40): 'none', 'r2' or 'r3' (the order of the
    total derivative whose mean square ``augment_dynamics`` integrates, latent_ode.py:82-108, 238-256).
    Closures, named as in init_model:
      gen_dynamics_wrap(x, t, params)            latent_ode.py:297   the count_nfe dynamics
      aug_dynamics((z, r), t, params)            augment_dynamics(gen_dynamics_wrap), latent_ode.py:238-256
    They are solved per trajectory (``jax.vmap(odeint)``, latent_ode.py:344-350) by ``easy_neural_ode_b200.odeint_vmap``;
    the element type of the state (float64 like the script, or float32) selects the kernel instantiation.
    Parameters: the Haiku dict {"gen_dynamics/linear[_i]": {"w", "b"}} or one flat vector in ravel order, per layer
    [b | W]."""

    def __init__(self, latent_dim, layers, units, reg="none"):
        if reg not in ["none"] + REGS:
            raise ValueError(f"--reg must be one of {['none'] + REGS}")
        if reg in REGS[2:]:
            raise NotImplementedError("Taylor orders above r3 are not built (BASELINE configs use r2/r3)")
        if not (1 <= latent_dim <= 64 and 1 <= units <= 64):
            raise ValueError("GenDynamics kernels hold widths up to 64")
        if not 0 <= layers <= 5:
            raise ValueError("GenDynamics supports gen_layers in 0..5")
        self.latent_dim, self.layers, self.units, self.reg = latent_dim, layers, units, reg
        self.dim = latent_dim
        self.reg_order = 0 if reg == "none" else REGS.index(reg) + 2
        self.dims = [latent_dim] + [units] * (layers + 1) + [latent_dim]
        self.gen_dynamics_wrap = EnoFunc(self, "plain", ("params",))
        self.dynamics_wrap = self.gen_dynamics_wrap
        self.aug_dynamics = EnoFunc(self, "aug", ("params",))

    def __repr__(self):
        return f"GenDynamics(latent_dim={self.latent_dim}, layers={self.layers}, units={self.units}, reg={self.reg!r})"

    @property
    def layer_names(self):
        return ["gen_dynamics/linear" + ("" if i == 0 else f"_{i}") for i in range(len(self.dims) - 1)]

    @property
    def n_params(self):
        return sum(i * o + o for i, o in zip(self.dims[:-1], self.dims[1:]))

    def desc(self, aug=_cabi.AUG_NONE, eps_in_args=False, reg_order=None):
        return _cabi.Desc(_cabi.ARCH_TANH_STACK, self.latent_dim, self.units,
                          self.reg_order if reg_order is None else reg_order, aug, 0, self.layers + 1)

    def init_params(self, seed=0, device="cpu", scale=1.0, dtype=torch.float64):
        """Haiku default init (GenDynamics is built without init_kwargs, latent_ode.py:293-296)."""
        g = torch.Generator().manual_seed(seed)

        def tn(shape, fan_in):
            w = torch.empty(shape, dtype=torch.float64)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            return (w * (scale / math.sqrt(fan_in))).to(dtype).to(device)

        return {n: {"w": tn((a, b), a), "b": torch.zeros(b, dtype=dtype, device=device)}
                for n, a, b in zip(self.layer_names, self.dims[:-1], self.dims[1:])}

    def flatten_params(self, params):
        if isinstance(params, torch.Tensor):
            if params.numel() != self.n_params:
                raise ValueError(f"flat params must have {self.n_params} elements")
            return params.reshape(-1)
        parts = []
        for n in self.layer_names:
            parts += [params[n]["b"].reshape(-1), params[n]["w"].reshape(-1)]
        return torch.cat(parts)

    def unflatten_params(self, flat):
        out, o = {}, 0
        for n, a, b in zip(self.layer_names, self.dims[:-1], self.dims[1:]):
            out[n] = {"b": flat[o:o + b], "w": flat[o + b:o + b + a * b].reshape(a, b)}
            o += b + a * b
        return out
