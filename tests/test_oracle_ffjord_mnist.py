"""The C4 oracle (oracle/ffjord_mnist_oracle.py: ConcatConv2D dynamics of ffjord_mnist.py) against independent restatements:
the closed-form Taylor term against nested forward-mode AD, the Hutchinson term against the exact trace on a tiny image,
the pre-flow's log-det against autograd, parameter count against SURVEY.md (P = 76,810)."""
import math

import torch

from oracle import ffjord_mnist_oracle as fm
from oracle.dynamics_oracle import sol_recursive


def test_param_count_and_zero_init():
    p = fm.init_conv_params()
    assert sum(v.numel() for l in p.values() for v in l.values()) == 76810       # C4, P = 76,810
    assert list(p) == sorted(p)                                                  # ravel_pytree order = call order
    x = torch.rand(2, 28, 28, 1)
    assert torch.count_nonzero(fm.conv_dynamics(p, x, 0.3)) == 0                 # last layer zero-initialised (214-215)


def test_closed_form_taylor_term_matches_nested_jvp():
    img = (6, 5, 1)
    p = fm.init_conv_params(hidden=8, seed=1, dtype=torch.float64, scale=1.5, last_scale=1.5)
    x = torch.randn(3, *img, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    f, y1 = fm.conv_taylor_closed(p, x, 0.4, img)
    f2, r = sol_recursive(lambda y, t: fm.conv_dynamics(p, y, t, img), x, 0.4, "r2")
    assert torch.allclose(f, f2, rtol=1e-12, atol=1e-12)
    assert torch.allclose(y1, r, rtol=1e-10, atol=1e-12)


def test_hutchinson_term_is_the_trace_in_expectation():
    img = (3, 3, 1)
    p = fm.init_conv_params(hidden=4, seed=3, dtype=torch.float64, scale=1.0, last_scale=1.0)
    x = torch.randn(1, *img, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    cl = fm.make_ffjord_mnist_closures("none", "our", img)
    J = torch.autograd.functional.jacobian(lambda y: fm.conv_dynamics(p, y, 0.2, img).reshape(-1), x.reshape(-1))
    n = 9
    tot = 0.0
    for k in range(2 ** n):          # every Rademacher vector: the average of eps^T J eps is exactly the trace
        eps = torch.tensor([1.0 if (k >> i) & 1 else -1.0 for i in range(n)], dtype=torch.float64).reshape(1, *img)
        tot += float(-cl["ffjord_dynamics"]((x, torch.zeros(1, 1)), 0.2, eps, p)[1])
    assert abs(tot / 2 ** n - float(torch.trace(J))) < 1e-10


def test_pre_flow_logdet_matches_autograd():
    x = torch.rand(2, 4, 4, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(7))
    y, logpy = fm.forward_pre_ode(x, torch.zeros(2, 1, dtype=torch.float64))
    J = torch.autograd.functional.jacobian(lambda v: fm.forward_pre_ode(v.reshape(1, 4, 4, 1), torch.zeros(1, 1))[0].reshape(-1),
                                           x[0].reshape(-1))
    # logpy = logpx - sum log dy/dx ... the script's sign convention: logdetgrad = -log(s - s^2) + log(1 - 2 alpha) = log dy/dx
    assert abs(float(logpy[0]) + float(torch.logdet(J))) < 1e-9
    assert math.isfinite(float(fm.bits_per_dim(y, -logpy)))
