"""GPU parity of the FFJORD tabular path (BASELINE config C2, ffjord_tabular.py) against the oracle closures and
against the golden split solve produced by the unmodified reference solver.  The contractions run on tcgen05
(3xTF32); tolerance: NFE / iteration counts exact, values rtol 1e-5 (north_star), single evaluations 2e-6-ish."""
import math

import numpy as np
import pytest
import torch

from tests.util import assert_agree, assert_count, exact_arith, load, rel_err

pytestmark = pytest.mark.gpu


def _to(p, dev, dtype=torch.float32):
    return {k: {kk: vv.to(device=dev, dtype=dtype) for kk, vv in v.items()} for k, v in p.items()}


def _oracle_flat(spec, tree):
    """oracle / reference gradient dict -> the library's flat order."""
    return spec.flatten_params(tree)


DEFAULT_TOL = 1.4e-8   # --rtol / --atol default of ffjord_tabular.py:34-35


def _agree(got, want, truth=None, what=""):
    assert_agree(got, want, truth, what)


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


def _truth_split(p, x, eps, gy, w, reg, reg_type):
    """fp64 oracle at 1e-10: the exact solution the fp32 solves approximate."""
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    d = torch.float64
    n_aux = len(w)
    cl = fo.make_ffjord_closures(reg, reg_type)
    p64 = _to(p, "cpu", d)
    z = torch.zeros(x.shape[0], 1, dtype=d)
    ts = torch.tensor([0., 1.], dtype=d)
    res, _ = oo.odeint_split_fwd(cl["fwd"], (x.to(d),) + (z,) * n_aux, ts, eps.to(d), p64, n_aux=n_aux, rtol=1e-10, atol=1e-10)
    gg = tuple(torch.zeros_like(r) for r in res)
    gg[0][-1] = gy.to(d)
    for q in range(n_aux):
        gg[1 + q][-1] = w[q]
    out = oo.odeint_split_rev(cl["aug_dynamics"], 1e-10, 1e-10, math.inf, res, ts, (eps.to(d), p64), gg)
    return dict(y=res[0], x0_bar=out[0][0], eps_bar=out[2], p_bar=out[3])


@pytest.mark.parametrize("name", ["css_sepaux_default.npz", "css_sepaux.npz"])
def test_sepaux_forward_and_adjoint_golden(cuda, name):
    """odeint_sepaux with the FFJORD closures against the golden vectors of the unmodified reference solver
    (tests/golden/make_golden.py:css_case), at the script-default tolerance and at 1e-5."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    g = load(name)
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
    if tol > 1e-7:
        assert float(nfe) == float(g["nfe"])
    else:
        # script-default tolerance: the count is set by arithmetic noise (tests/util.py:exact_arith) - the unmodified
        # reference solver with exactly evaluated dynamics gives the other end of the admissible band
        cl = fo.make_ffjord_closures("r2", "our")
        zz = torch.zeros(B, 1)
        _, nfe_exact = oo.odeint_split_fwd(exact_arith(cl["fwd"]), (torch.as_tensor(g["x0"]), zz, zz), torch.tensor([0., 1.]),
                                           torch.as_tensor(g["eps"]), p, n_aux=2, rtol=tol, atol=tol)
        assert_count(float(nfe), float(g["nfe"]), float(nfe_exact), "nfe")
    assert tuple(dl.shape) == (2, B, 1) and float(dl.detach().abs().max()) == 0.0
    loss = (ys[-1] * torch.as_tensor(g["gy"], device=cuda)).sum() + dl[-1].sum() / B + 0.01 * rs[-1].sum() / B
    loss.backward()
    sort_flat = lambda tree: torch.cat([v.reshape(-1) for n in spec.layer_names for _, v in sorted(tree[n].items())])
    tr = None
    if tol > 1e-7:
        t = _truth_split(p, torch.as_tensor(g["x0"]), torch.as_tensor(g["eps"]), torch.as_tensor(g["gy"]), [1.0 / B, 0.01 / B],
                         "r2", "our")
        tr = dict(y=t["y"], x0_bar=t["x0_bar"], eps_bar=t["eps_bar"], p_bar=sort_flat(t["p_bar"]))
    _agree(ys, g["y"], tr and tr["y"], "y")
    _agree(x0.grad, g["x0_bar"], tr and tr["x0_bar"], "x0_bar")
    _agree(eps.grad, g["eps_bar"], tr and tr["eps_bar"], "eps_bar")
    # golden params_bar is in ravel_pytree order of the oracle's per-layer dicts (sorted keys)
    _agree(sort_flat(spec.unflatten_params(pflat.grad.cpu())), g["params_bar"], tr and tr["p_bar"], "params_bar")


@pytest.mark.parametrize("tol", [DEFAULT_TOL, 1e-5])
@pytest.mark.parametrize("reg_type", ["our", "fin"])
def test_split_solves_match_oracle(cuda, reg_type, tol):
    """odeint_sepaux (reg_type 'our') and odeint_fin_sepaux (reg_type 'fin', three auxiliary states: lib/ode.py:699-718)
    at a mid size against the oracle: iteration counts exact, values and gradients per _agree."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    B, D, H = 64, 43, 128
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
    if tol < 1e-7:
        # below fp32 epsilon the counts are a property of the arithmetic noise (tests/util.py:exact_arith): the band spanned
        # by the fp32 oracle and the exact-arithmetic oracle (here 50 / 44 NFE and 11 / 9 adjoint iterations)
        resx, nfe_x = oo.odeint_split_fwd(exact_arith(cl["fwd"]), (x,) + (z,) * n_aux, ts, eps, p, n_aux=n_aux, rtol=tol, atol=tol)
        bsx = []
        oo.odeint_split_rev(exact_arith(cl["aug_dynamics"]), tol, tol, math.inf, resx, ts, (eps, p), gg, bsx)
        assert_count(float(nfe), float(nfe_w), float(nfe_x), "nfe")
        assert_count(eno.last_stats["bwd"][0]["n_iters"], bst[0]["n_iters"], bsx[0]["n_iters"], "adjoint iterations")
    else:
        assert float(nfe) == float(nfe_w)
        assert eno.last_stats["bwd"][0]["n_iters"] == bst[0]["n_iters"]
    tr = None
    if tol > 1e-7:
        t = _truth_split(p, x, eps, gy, w, reg, reg_type)
        tr = dict(y=t["y"], x0_bar=t["x0_bar"], eps_bar=t["eps_bar"], p_bar=_oracle_flat(spec, t["p_bar"]))
    _agree(outs[0], res[0], tr and tr["y"], "y")
    _agree(xg.grad, want[0][0], tr and tr["x0_bar"], "x0_bar")
    _agree(eg.grad, want[2], tr and tr["eps_bar"], "eps_bar")
    _agree(pflat.grad, _oracle_flat(spec, want[3]), tr and tr["p_bar"], "params_bar")


def test_ffjord_and_all_aug_forward_solves(cuda):
    """The evaluation-path solves: odeint on ffjord_dynamics (the NFE count of nfe_fn, ffjord_tabular.py:329-347) and on
    all_aug_dynamics (forward_all, 286-297) against the oracle."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    B, D, H, tol = 32, 43, 96, DEFAULT_TOL
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


@pytest.mark.parametrize("tol", [1e-6, DEFAULT_TOL])
def test_c2_shape_train_path(cuda, tol):
    """BASELINE configs[1] shape (B = 1000, D = 43, hidden 860 x 2, --reg r2; ffjord_tabular.py:29-58).  At 1e-6 (error
    estimates above the fp32 round-off floor): forward NFE and adjoint iteration count identical to the oracle.  At the
    script-default 1.4e-8 the counts are set by arithmetic noise and must lie in the band spanned by the fp32 oracle and
    the exact-arithmetic oracle (tests/util.py:exact_arith).  Values and gradients 1e-5 in both."""
    import os
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    B, D, H = 1000, 43, 860
    p, x, eps = _case(B, D, H, seed=0, scale=1.0)
    spec = eno.ConcatSquashDynamics(D, H, reg="r2")
    cl = fo.make_ffjord_closures("r2", "our")
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))

    def oracle_run(fwd, rev):
        res, nfe_w = oo.odeint_split_fwd(fwd, (x, z, z), ts, eps, p, n_aux=2, rtol=tol, atol=tol)
        gg = tuple(torch.zeros_like(r) for r in res)
        gg[0][-1], gg[1][-1], gg[2][-1] = res[0][-1] / B, 1.0 / B, 0.01 / B    # the loss's cotangents (ffjord_tabular.py:402-442)
        bst = []
        want = oo.odeint_split_rev(rev, tol, tol, math.inf, res, ts, (eps, p), gg, bst)
        return res, float(nfe_w), gg, bst[0]["n_iters"], want

    res, nfe_w, gg, bwd_w, want = oracle_run(cl["fwd"], cl["aug_dynamics"])
    tr = None
    if tol < 1e-7:
        # the exact-arithmetic solve at this tolerance doubles as the third party of assert_agree (error ~ 1e-8): the fp32
        # oracle's own step sequence depends on the host's GEMM threading, and two fp32 solves that take different steps
        # need not agree to 1e-5 in every gradient entry
        res_x, nfe_x, _, bwd_x, want_x = oracle_run(exact_arith(cl["fwd"]), exact_arith(cl["aug_dynamics"]))
        tr = dict(y=res_x[0], x0_bar=want_x[0][0], p_bar=spec.flatten_params(want_x[3]))
    else:
        nfe_x, bwd_x = nfe_w, bwd_w
    xg = x.to(cuda).requires_grad_(True)
    pflat = spec.flatten_params(_to(p, cuda)).clone().requires_grad_(True)
    zc = z.to(cuda)
    (ys, dl, rs), nfe = eno.odeint_sepaux(spec.fwd, spec.aug_dynamics, (xg, zc, zc), ts, eps.to(cuda), pflat, rtol=tol, atol=tol)
    loss = (ys[-1] * gg[0][-1].to(cuda)).sum() + dl[-1].sum() / B + 0.01 * rs[-1].sum() / B
    loss.backward()
    print(f"C2 tol {tol}: nfe gpu {float(nfe)} oracle fp32 {nfe_w} exact-arithmetic {nfe_x}; adjoint iterations gpu "
          f"{eno.last_stats['bwd'][0]['n_iters']} oracle fp32 {bwd_w} exact-arithmetic {bwd_x}")
    slack = 1 if tol < 1e-7 else 0    # two samples of the round-off noise do not bound a third (tests/util.py:assert_count)
    assert_count(float(nfe), nfe_w, nfe_x, "nfe", slack=6 * slack)
    assert_count(eno.last_stats["bwd"][0]["n_iters"], bwd_w, bwd_x, "adjoint iterations", slack=slack)
    if tol > 1e-7:   # fp64 oracle at 1e-10 as third party (tests/util.py:assert_agree)
        t = _truth_split(p, x, eps, gg[0][-1], [1.0 / B, 0.01 / B], "r2", "our")
        tr = dict(y=t["y"], x0_bar=t["x0_bar"], p_bar=_oracle_flat(spec, t["p_bar"]))
    _agree(ys, res[0], tr and tr["y"], "y")
    _agree(xg.grad, want[0][0], tr and tr["x0_bar"], "x0_bar")
    _agree(pflat.grad, _oracle_flat(spec, want[3]), tr and tr["p_bar"], "params_bar")
