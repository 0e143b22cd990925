"""GPU parity of the FFJORD tabular path (BASELINE config C2, ffjord_tabular.py) against the oracle closures and
against the golden split solve produced by the unmodified reference solver.  The contractions run on tcgen05
(3xTF32); tolerance: NFE / iteration counts exact, values rtol 1e-5 (north_star), single evaluations 2e-6-ish."""
import math

import numpy as np
import pytest
import torch

from tests.util import load, rel_err

pytestmark = pytest.mark.gpu


def _to(p, dev, dtype=torch.float32):
    return {k: {kk: vv.to(device=dev, dtype=dtype) for kk, vv in v.items()} for k, v in p.items()}


def _oracle_flat(spec, tree):
    """oracle / reference gradient dict -> the library's flat order."""
    return spec.flatten_params(tree)


def _case(B, D, H, seed=3, scale=1.5, layers=2):
    from oracle import ffjord_oracle as fo
    p = fo.init_css_params(D, H, n_hidden_layers=layers, seed=seed, scale=scale)
    g = torch.Generator().manual_seed(9)
    x = torch.randn((B, D), generator=g)
    eps = torch.randn((B, D), generator=g)
    return p, x, eps


@pytest.mark.parametrize("shape", [(5, 6, 12), (37, 43, 100), (130, 20, 200)])
def test_plain_rhs_matches_oracle(cuda, shape):
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo
    B, D, H = shape
    p, x, _ = _case(B, D, H)
    spec = eno.ConcatSquashDynamics(D, H)
    got = eno.rhs_closure(spec.dynamics_wrap, 0, x.to(cuda), 0.3, _to(p, cuda))
    want = fo.css_dynamics(_to(p, "cpu", torch.float64), x.double(), 0.3)
    assert rel_err(got, want) < 3e-6


@pytest.mark.parametrize("shape", [(5, 6, 12), (37, 43, 100)])
def test_all_aug_rhs_matches_oracle(cuda, shape):
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo
    B, D, H = shape
    p, x, eps = _case(B, D, H)
    spec = eno.ConcatSquashDynamics(D, H, reg="r2")
    dy, aux = eno.rhs_closure(spec.all_aug_dynamics, 1, x.to(cuda), 0.7, eps.to(cuda), _to(p, cuda))
    z = torch.zeros(B, 1, dtype=torch.float64)
    want = fo.make_ffjord_closures("r2", "our")["all_aug_dynamics"]((x.double(), z, z, z, z), 0.7, eps.double(),
                                                                     _to(p, "cpu", torch.float64))
    assert rel_err(dy, want[0]) < 3e-6
    for q in range(4):
        assert rel_err(aux[q], want[1 + q].reshape(-1)) < 1e-5, q


@pytest.mark.parametrize("reg_type,reg", [("our", "r2"), ("our", "none"), ("fin", "none")])
@pytest.mark.parametrize("shape", [(5, 6, 12), (37, 43, 100)])
def test_adjoint_rhs_matches_torch_func_vjp(cuda, shape, reg_type, reg):
    """The hand-derived reverse sweep (three row streams through the GEMM pipeline) against torch.func.vjp of the
    oracle closure in fp64."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo
    B, D, H = shape
    p, x, eps = _case(B, D, H)
    spec = eno.ConcatSquashDynamics(D, H, reg=reg, reg_type=reg_type)
    n_aux = 2 if reg_type == "our" else 3
    g = torch.Generator().manual_seed(21)
    yb = torch.randn((B, D), generator=g)
    ab = torch.randn((n_aux, B), generator=g)
    got = eno.rhs_closure(spec.aug_dynamics, 2, x.to(cuda), 0.4, eps.to(cuda), _to(p, cuda), y_bar=yb.to(cuda), aux_bar=ab.to(cuda))
    cl = fo.make_ffjord_closures(reg, reg_type)["aug_dynamics"]
    p64 = _to(p, "cpu", torch.float64)
    z = torch.zeros(B, 1, dtype=torch.float64)

    def f(y, t, e, pr):
        out = cl((y,) + (z,) * n_aux, t, e, pr)
        return (out[0],) + tuple(o.reshape(-1) for o in out[1:])
    out, vjp = torch.func.vjp(f, x.double(), torch.tensor(0.4, dtype=torch.float64), eps.double(), p64)
    cot = (yb.double(),) + tuple(ab[q].double() for q in range(n_aux))
    vy, vt, ve, vp = vjp(cot)
    assert rel_err(got[0], out[0]) < 3e-6
    for q in range(n_aux):
        assert rel_err(got[1][q], out[1 + q]) < 1e-5
    assert rel_err(got[2], vy) < 1e-5
    assert abs(float(got[3]) - float(vt)) < 1e-5 * max(1.0, abs(float(vt)))
    assert rel_err(got[4], _oracle_flat(spec, vp)) < 1e-5
    assert rel_err(got[5], ve) < 1e-5


def test_sepaux_forward_and_adjoint_golden(cuda):
    """odeint_sepaux with the FFJORD closures against the golden vectors of the unmodified reference solver
    (tests/golden/make_golden.py:css_case)."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo
    g = load("css_sepaux.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    p = fo.init_css_params(D, H, seed=int(g["seed"]), scale=float(g["scale"]))
    spec = eno.ConcatSquashDynamics(D, H, reg="r2")
    tol = float(g["tol"])
    x0 = torch.as_tensor(g["x0"], device=cuda).requires_grad_(True)
    eps = torch.as_tensor(g["eps"], device=cuda).requires_grad_(True)
    pflat = spec.flatten_params(_to(p, cuda)).clone().requires_grad_(True)
    z = torch.zeros(B, 1, device=cuda)
    (ys, dl, rs), nfe = eno.odeint_sepaux(spec.fwd, spec.aug_dynamics, (x0, z, z), torch.tensor([0., 1.]), eps, pflat,
                                          rtol=tol, atol=tol)
    assert float(nfe) == float(g["nfe"])
    assert tuple(dl.shape) == (2, B, 1) and float(dl.abs().max()) == 0.0
    assert rel_err(ys, g["y"]) < 1e-5
    loss = (ys[-1] * torch.as_tensor(g["gy"], device=cuda)).sum() + dl[-1].sum() / B + 0.01 * rs[-1].sum() / B
    loss.backward()
    assert rel_err(x0.grad, g["x0_bar"]) < 1e-5
    assert rel_err(eps.grad, g["eps_bar"]) < 1e-5
    # golden params_bar is in ravel_pytree order of the oracle's per-layer dicts (sorted keys)
    tree = spec.unflatten_params(pflat.grad.cpu())
    flat = torch.cat([v.reshape(-1) for n in spec.layer_names for _, v in sorted(tree[n].items())])
    assert rel_err(flat, g["params_bar"]) < 1e-5


@pytest.mark.parametrize("reg_type", ["our", "fin"])
def test_split_solves_match_oracle(cuda, reg_type):
    """odeint_sepaux (reg_type 'our') and odeint_fin_sepaux (reg_type 'fin', three auxiliary states: lib/ode.py:699-718)
    at a mid size against the oracle: iteration counts exact, gradients 1e-5."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    B, D, H, tol = 64, 43, 128, 1e-5
    p, x, eps = _case(B, D, H, scale=1.3)
    reg = "r2" if reg_type == "our" else "none"
    spec = eno.ConcatSquashDynamics(D, H, reg=reg, reg_type=reg_type)
    n_aux = 2 if reg_type == "our" else 3
    cl = fo.make_ffjord_closures(reg, reg_type)
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    res, nfe_w, fst = oo.odeint_split_fwd(cl["fwd"], (x,) + (z,) * n_aux, ts, eps, p, n_aux=n_aux, rtol=tol, atol=tol,
                                          return_stats=True)
    gy = torch.randn((B, D), generator=torch.Generator().manual_seed(4))
    w = [1.0 / B, 0.01 / B, 0.02 / B][:n_aux]
    gg = tuple(torch.zeros_like(r) for r in res)
    gg[0][-1] = gy
    for q in range(n_aux):
        gg[1 + q][-1] = w[q]
    bst = []
    want = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, p), gg, bst)

    xg = x.to(cuda).requires_grad_(True)
    eg = eps.to(cuda).requires_grad_(True)
    pflat = spec.flatten_params(_to(p, cuda)).clone().requires_grad_(True)
    zc = z.to(cuda)
    fn = eno.odeint_sepaux if reg_type == "our" else eno.odeint_fin_sepaux
    outs, nfe = fn(spec.fwd, spec.aug_dynamics, (xg,) + (zc,) * n_aux, ts, eg, pflat, rtol=tol, atol=tol)
    loss = (outs[0][-1] * gy.to(cuda)).sum() + sum(w[q] * outs[1 + q][-1].sum() for q in range(n_aux))
    loss.backward()
    assert float(nfe) == float(nfe_w)
    assert eno.last_stats["bwd"][0]["n_iters"] == bst[0]["n_iters"]
    assert rel_err(outs[0], res[0]) < 1e-5
    assert rel_err(xg.grad, want[0][0]) < 1e-5
    assert rel_err(eg.grad, want[2]) < 1e-5
    assert rel_err(pflat.grad, _oracle_flat(spec, want[3])) < 1e-5


def test_ffjord_and_all_aug_forward_solves(cuda):
    """The evaluation-path solves: odeint on ffjord_dynamics (the NFE count of nfe_fn, ffjord_tabular.py:329-347) and on
    all_aug_dynamics (forward_all, 286-297) against the oracle."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    B, D, H, tol = 32, 43, 96, 1e-6
    p, x, eps = _case(B, D, H, scale=1.3)
    spec = eno.ConcatSquashDynamics(D, H, reg="r2")
    cl = fo.make_ffjord_closures("r2", "our")
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    (y_w, p_w), nfe_w = oo.odeint(cl["ffjord_dynamics"], (x, z), ts, eps, p, rtol=tol, atol=tol)
    (y_g, p_g), nfe_g = eno.odeint(spec.ffjord_dynamics, (x.to(cuda), z.to(cuda)), ts, eps.to(cuda), _to(p, cuda), rtol=tol, atol=tol)
    assert float(nfe_g) == float(nfe_w)
    assert rel_err(y_g, y_w) < 1e-5 and rel_err(p_g[-1], p_w[-1]) < 1e-5
    z1 = torch.zeros(B)
    want, nfe_w = oo.odeint(cl["all_aug_dynamics"], (x, z, z1, z1, z1), ts, eps, p, rtol=tol, atol=tol)
    got, nfe_g = eno.odeint(spec.all_aug_dynamics, (x.to(cuda), z.to(cuda), z1.to(cuda), z1.to(cuda), z1.to(cuda)), ts,
                            eps.to(cuda), _to(p, cuda), rtol=tol, atol=tol)
    assert float(nfe_g) == float(nfe_w)
    for a, b in zip(got, want):
        assert rel_err(a[-1], b[-1]) < 1e-5


def test_c2_shape_train_path(cuda):
    """BASELINE configs[1] shape (B = 1000, D = 43, hidden 860 x 2, --reg r2, ffjord_tabular.py defaults) at tol 1e-5:
    forward NFE and adjoint iteration count identical to the oracle, gradients 1e-5."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    B, D, H, tol = 1000, 43, 860, 1e-5
    p, x, eps = _case(B, D, H, seed=0, scale=1.0)
    spec = eno.ConcatSquashDynamics(D, H, reg="r2")
    cl = fo.make_ffjord_closures("r2", "our")
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    torch.set_num_threads(max(1, len(__import__("os").sched_getaffinity(0))))
    res, nfe_w = oo.odeint_split_fwd(cl["fwd"], (x, z, z), ts, eps, p, n_aux=2, rtol=tol, atol=tol)
    gg = tuple(torch.zeros_like(r) for r in res)
    gg[0][-1], gg[1][-1], gg[2][-1] = res[0][-1] / B, 1.0 / B, 0.01 / B      # the loss's cotangents (ffjord_tabular.py:402-442)
    bst = []
    want = oo.odeint_split_rev(cl["aug_dynamics"], tol, tol, math.inf, res, ts, (eps, p), gg, bst)
    xg = x.to(cuda).requires_grad_(True)
    pflat = spec.flatten_params(_to(p, cuda)).clone().requires_grad_(True)
    zc = z.to(cuda)
    (ys, dl, rs), nfe = eno.odeint_sepaux(spec.fwd, spec.aug_dynamics, (xg, zc, zc), ts, eps.to(cuda), pflat, rtol=tol, atol=tol)
    loss = (ys[-1] * gg[0][-1].to(cuda)).sum() + dl[-1].sum() / B + 0.01 * rs[-1].sum() / B
    loss.backward()
    print("C2 fwd", eno.last_stats["fwd"], "bwd", eno.last_stats["bwd"][0], "oracle bwd", bst[0])
    assert float(nfe) == float(nfe_w)
    assert eno.last_stats["bwd"][0]["n_iters"] == bst[0]["n_iters"]
    assert rel_err(ys, res[0]) < 1e-5
    assert rel_err(xg.grad, want[0][0]) < 1e-5
    assert rel_err(pflat.grad, _oracle_flat(spec, want[3])) < 1e-5
