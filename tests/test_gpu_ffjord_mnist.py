"""BASELINE config C4 (ffjord_mnist.py), forward closures: the ConcatConv2D dynamics on the CUDA path (csrc/ccv_kernels.cuh:
direct first / last layers, the two C -> C layers as 9-tap loops of the tcgen05 3xTF32 kernel over a zero-padded position
grid) against the oracle (oracle/ffjord_mnist_oracle.py) and against golden solves of the unmodified reference solver
(tests/golden/conv_ffjord.npz, make_conv.py).  The adjoint of the convolutional stack is not built: asserted to raise."""
import math

import pytest
import torch

from tests.util import assert_agree, assert_count, exact_arith, load, rel_err

pytestmark = pytest.mark.gpu


def _close(got, want, tol, what=""):
    """The solver's own mixed tolerance, with the usual factor for two fp32 solves of the same problem whose step sizes
    differ after the round-off-driven first error estimate: |got - want| <= 3 tol (1 + max |want|)."""
    got = torch.as_tensor(got).double().cpu().reshape(-1)
    want = torch.as_tensor(want).double().cpu().reshape(-1)
    err = float((got - want).abs().max())
    assert err <= 3 * tol * (1.0 + float(want.abs().max())), (what, err)


def _case(B, image, hidden, seed=0, scale=1.5):
    from oracle import ffjord_mnist_oracle as fm
    p = fm.init_conv_params(hidden=hidden, seed=seed, scale=scale, last_scale=scale)
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.randn((B,) + image + (1,), generator=g)
    eps = torch.randint(0, 2, (B,) + image + (1,), generator=g).float() * 2 - 1
    return p, x, eps


def _to(p, dev):
    return {k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p.items()}


@pytest.mark.parametrize("shape", [(3, (6, 6), 32), (2, (28, 28), 64), (5, (10, 6), 64)])
def test_conv_rhs_matches_oracle(cuda, shape):
    """One evaluation: f, and the all-aug closure (dy, -div, r_2, dfro, dkin) - every border class of the time channel,
    both tangent streams, halo handling of the padded grid."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_mnist_oracle as fm
    B, image, hidden = shape
    p, x, eps = _case(B, image, hidden)
    spec = eno.ConcatConvDynamics(image, hidden, reg="r2")
    cl = fm.make_ffjord_mnist_closures("r2", "our", image + (1,))
    t = 0.37
    want = fm.conv_dynamics(p, x, t, image + (1,))
    got = eno.rhs_closure(spec.dynamics_wrap, 0, x.to(cuda), t, _to(p, cuda))
    assert rel_err(got.reshape(-1), want.reshape(-1)) < 1e-5
    z = torch.zeros(B, 1)
    wa = cl["all_aug_dynamics"]((x, z, z, z, z), t, eps, p)
    dy, aux = eno.rhs_closure(spec.all_aug_dynamics, 1, x.to(cuda), t, eps.to(cuda), _to(p, cuda))
    assert rel_err(dy.reshape(-1), wa[0].reshape(-1)) < 1e-5
    for q, name in enumerate(("-div", "r2", "fro", "kin")):
        assert rel_err(aux[q], wa[1 + q].reshape(-1)) < 2e-5, name
    # nq = 1 path (no Taylor stream): ffjord_dynamics
    wf = cl["ffjord_dynamics"]((x, z), t, eps, p)
    dy, aux = eno.rhs_closure(spec.ffjord_dynamics, 1, x.to(cuda), t, eps.to(cuda), _to(p, cuda))
    assert rel_err(dy.reshape(-1), wf[0].reshape(-1)) < 1e-5
    assert rel_err(aux[0], wf[1].reshape(-1)) < 2e-5


def test_zero_initialised_last_layer(cuda):
    """The script's initial parameters (last layer's w = 0, ffjord_mnist.py:214-215): f == 0, every solve takes the
    minimum number of steps."""
    import easy_neural_ode_b200 as eno
    spec = eno.ConcatConvDynamics((28, 28), 64)
    p = spec.init_params(seed=0, device=cuda)
    assert spec.flatten_params(p).numel() == 76810 == spec.n_params
    x = torch.rand(4, 28, 28, 1, device=cuda)
    ys, nfe = eno.odeint(spec.dynamics_wrap, x, torch.tensor([0., 1.]), p)
    assert torch.equal(ys[-1], x)


def test_golden_reference_solves(cuda):
    """nfe_fn's solve (ffjord_dynamics on (y, delta_logp)), forward_all's solve (all_aug_dynamics) and the plain solve
    against the unmodified reference solver: NFE exact; the golden solves run at tol = 1e-5, where two fp32 implementations
    agree to the solve's own accuracy (bar 3 tol, tests/README.md)."""
    import easy_neural_ode_b200 as eno
    g = load("conv_ffjord.npz")
    B, img, hidden, tol = int(g["B"]), tuple(int(v) for v in g["img"]), int(g["hidden"]), float(g["tol"])
    x0, eps = torch.as_tensor(g["x0"]).to(cuda), torch.as_tensor(g["eps"]).to(cuda)
    pflat = torch.as_tensor(g["params"]).to(cuda)
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1, device=cuda)
    bar = max(1e-5, 3 * tol)
    # float64 oracle at 1e-10 as third party for the small, cancellation-prone auxiliary outputs (tests/util.py:assert_agree)
    from oracle import ffjord_mnist_oracle as fm, ode_oracle as oo
    spec0 = eno.ConcatConvDynamics(img[:2], hidden)
    p64 = {k: {kk: vv.double().cpu() for kk, vv in v.items()} for k, v in spec0.unflatten_params(pflat).items()}
    cl64 = fm.make_ffjord_mnist_closures("r2", "our", img)
    z64 = torch.zeros(B, 1, dtype=torch.float64)
    truth, _ = oo.odeint(cl64["all_aug_dynamics"], (x0.double().cpu(), z64, z64, z64, z64), ts.double(), eps.double().cpu(), p64,
                         rtol=1e-10, atol=1e-10)
    ys, nfe = eno.odeint(spec0.dynamics_wrap, x0, ts, pflat, rtol=tol, atol=tol)
    assert float(nfe) == float(g["plain_nfe"])
    assert rel_err(ys, g["plain_y"]) < bar
    for reg in ("none", "r2"):
        spec = eno.ConcatConvDynamics(img[:2], hidden, reg=reg)
        (ys, dl), nfe = eno.odeint(spec.ffjord_dynamics, (x0, z), ts, eps, pflat, rtol=tol, atol=tol)
        assert float(nfe) == float(g[f"ffjord_{reg}_nfe"])
        assert rel_err(ys, g[f"ffjord_{reg}_y"]) < bar
        assert_agree(dl[-1], g[f"ffjord_{reg}_dlogp"][-1], truth[1][-1], "delta_logp", rtol=bar)
    spec = eno.ConcatConvDynamics(img[:2], hidden, reg="r2")
    (ys, dl, rs, fro, kin), nfe = eno.odeint(spec.all_aug_dynamics, (x0, z, z, z, z), ts, eps, pflat, rtol=tol, atol=tol)
    assert float(nfe) == float(g["all_nfe"])
    assert rel_err(ys, g["all_y"]) < bar
    for q, (got, key) in enumerate(((dl, "all_dlogp"), (rs, "all_r"), (fro, "all_fro"), (kin, "all_kin"))):
        assert_agree(got[-1], g[key][-1], truth[1 + q][-1], key, rtol=bar)


@pytest.mark.parametrize("tol", [1e-5, 1e-6])
def test_script_size_forward_solves_vs_oracle(cuda, tol):
    """28 x 28 images, 64 hidden channels (the script's network), a shard of 4 images: the log-density solve
    (ffjord_dynamics) against the oracle.  NFE: equal to the fp32 oracle where the exact-arithmetic oracle agrees with it,
    inside the band the two span otherwise (the first, tiny step's error estimate is round-off and steers the second step
    size: tests/util.py:exact_arith); y(t1) and delta_logp within 1e-5 of the oracle whose step count it shares."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_mnist_oracle as fm, ode_oracle as oo
    B, image, hidden = 4, (28, 28), 64
    p, x, eps = _case(B, image, hidden, seed=3, scale=1.2)
    spec = eno.ConcatConvDynamics(image, hidden)
    cl = fm.make_ffjord_mnist_closures("none", "our", image + (1,))
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    want, nfe_w = oo.odeint(cl["ffjord_dynamics"], (x, z), ts, eps, p, rtol=tol, atol=tol)
    want_x, nfe_x = oo.odeint(exact_arith(cl["ffjord_dynamics"]), (x, z), ts, eps, p, rtol=tol, atol=tol)
    (ys, dl), nfe = eno.odeint(spec.ffjord_dynamics, (x.to(cuda), z.to(cuda)), ts, eps.to(cuda), _to(p, cuda), rtol=tol, atol=tol)
    print(f"C4 ffjord_dynamics tol {tol:g}: nfe cuda {float(nfe):.0f}, fp32 oracle {float(nfe_w):.0f}, exact-arithmetic oracle {float(nfe_x):.0f}")
    assert_count(float(nfe), float(nfe_w), float(nfe_x), "nfe")
    # values against the oracle whose step sequence this is (the two oracles are different discretisations of the problem)
    ref = want if float(nfe) == float(nfe_w) else want_x
    assert float(nfe) in (float(nfe_w), float(nfe_x))
    assert rel_err(ys[-1], ref[0][-1]) < 1e-5
    assert rel_err(dl[-1], ref[1][-1]) < 1e-5


def test_pre_flow_and_bits_per_dim(cuda):
    """forward_all of the script around the solve: logit pre-flow, all_aug solve, bits/dim (ffjord_mnist.py:155-168,
    409-417, 455-470) through easy_neural_ode_b200.ffjord_mnist_step against the oracle glue."""
    import easy_neural_ode_b200 as eno
    from easy_neural_ode_b200 import ffjord_mnist_step as fs
    from oracle import ffjord_mnist_oracle as fm, ode_oracle as oo
    B, image, hidden, tol = 4, (28, 28), 64, 1e-5
    p, _, eps = _case(B, image, hidden, seed=5, scale=1.0)
    images = torch.rand((B,) + image + (1,), generator=torch.Generator().manual_seed(9))
    model = fs.ImageFfjord(image, hidden, reg="r2", rtol=tol, atol=tol, device=cuda)
    model.params = model.spec.flatten_params(_to(p, cuda)).clone()
    out = model.forward_all(images.to(cuda), eps.to(cuda))
    cl = fm.make_ffjord_mnist_closures("r2", "our", image + (1,))
    y, logpy = fm.forward_pre_ode(images, torch.zeros(B, 1))
    z = torch.zeros(B, 1)
    (ys, dl, rs, fro, kin), nfe_w = oo.odeint(cl["all_aug_dynamics"], (y, logpy, z, z, z), torch.tensor([0., 1.]), eps, p,
                                            rtol=tol, atol=tol)
    assert out["nfe"] == float(nfe_w)
    assert rel_err(out["z"], ys[-1]) < 3e-5
    assert rel_err(out["delta_logp"], dl[-1]) < 3e-5
    assert abs(out["bits_per_dim"] - float(fm.bits_per_dim(ys[-1], dl[-1]))) < 1e-4
    _close(out["r"], rs[-1], tol, "r")
    _close(out["fro"], fro[-1], tol, "fro")
    _close(out["kin"], kin[-1], tol, "kin")


@pytest.mark.parametrize("reg_type,reg", [("our", "r2"), ("our", "none"), ("fin", "none")])
@pytest.mark.parametrize("shape", [(3, (6, 6), 32), (2, (28, 28), 64)])
def test_adjoint_rhs_matches_torch_func_vjp(cuda, shape, reg_type, reg):
    """The adjoint's integrand (lib/ode.py:1498-1504): the pull-back of (y_bar, aux_bar) through aug_dynamics at (y, t)
    wrt y, t, eps and every parameter, against torch.func.vjp of the oracle closure in float64."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_mnist_oracle as fm
    B, image, hidden = shape
    p, x, eps = _case(B, image, hidden, seed=7)
    img = image + (1,)
    spec = eno.ConcatConvDynamics(image, hidden, reg=reg, reg_type=reg_type)
    cl = fm.make_ffjord_mnist_closures(reg, reg_type, img)
    n_aux = 2 if reg_type == "our" else 3
    g = torch.Generator().manual_seed(31)
    y_bar = torch.randn((B,) + img, generator=g)
    aux_bar = torch.randn((n_aux, B), generator=g)
    t = 0.41
    p64 = {k: {kk: vv.double() for kk, vv in v.items()} for k, v in p.items()}
    z = torch.zeros(B, 1, dtype=torch.float64)

    def f(y, tt, e, pp):
        out = cl["aug_dynamics"]((y, z) + (z,) * (n_aux - 1), tt, e, pp)
        return (out[0],) + tuple(o.reshape(-1) for o in out[1:])

    outs, vjp = torch.func.vjp(f, x.double(), torch.tensor(t, dtype=torch.float64), eps.double(), p64)
    cot = (y_bar.double(),) + tuple(aux_bar[q].double() for q in range(n_aux))
    wy, wt, we, wp = vjp(cot)
    dy, aux, zb, tbar, pbar, eb = eno.rhs_closure(spec.aug_dynamics, 2, x.to(cuda), t, eps.to(cuda), _to(p, cuda), y_bar=y_bar.to(cuda),
                                                  aux_bar=aux_bar.to(cuda))
    assert rel_err(dy.reshape(-1), outs[0].reshape(-1)) < 1e-5
    for q in range(n_aux):
        assert rel_err(aux[q], outs[1 + q]) < 2e-5, ("aux", q)
    assert rel_err(zb.reshape(-1), wy.reshape(-1)) < 2e-5
    assert rel_err(eb.reshape(-1), we.reshape(-1)) < 2e-5
    assert abs(float(tbar) - float(wt)) <= 2e-5 * max(1.0, abs(float(wt)))
    want_p = spec.flatten_params(wp)
    assert rel_err(pbar, want_p) < 2e-5
    for name, lo, hi in _param_blocks(spec):        # every block on its own scale (the last layer's are much larger)
        assert rel_err(pbar[lo:hi], want_p[lo:hi]) < 5e-5, name


def _param_blocks(spec):
    o = 0
    for n, c_in, c_out in zip(spec.layer_names, spec.chans[:-1], spec.chans[1:]):
        yield n + "/b", o, o + c_out
        o += c_out
        yield n + "/w", o, o + 9 * (c_in + 1) * c_out
        o += 9 * (c_in + 1) * c_out


@pytest.mark.parametrize("reg_type", ["our", "fin"])
def test_split_solve_and_adjoint_vs_oracle(cuda, reg_type):
    """The training call of ffjord_mnist.py (340-347: odeint_sepaux / odeint_fin_sepaux with the conv closures) on 12 x 12
    images, 32 channels: forward NFE and adjoint iteration count in the band of the fp32 / exact-arithmetic oracles
    (tests/util.py), gradients wrt the image, eps and every parameter against the oracle whose step counts it shares."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_mnist_oracle as fm, ode_oracle as oo
    B, image, hidden, tol = 3, (12, 12), 32, 1e-6
    img = image + (1,)
    p, x, eps = _case(B, image, hidden, seed=11, scale=1.3)
    reg = "r2" if reg_type == "our" else "none"
    spec = eno.ConcatConvDynamics(image, hidden, reg=reg, reg_type=reg_type)
    n_aux = 2 if reg_type == "our" else 3
    cl = fm.make_ffjord_mnist_closures(reg, reg_type, img)
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    gy = torch.randn((B,) + img, generator=torch.Generator().manual_seed(4))
    w = [1.0 / B, 0.01 / B, 0.02 / B][:n_aux]

    def oracle_run(fwd, rev):
        res, nfe_w = oo.odeint_split_fwd(fwd, (x,) + (z,) * n_aux, ts, eps, p, n_aux=n_aux, rtol=tol, atol=tol)
        gg = tuple(torch.zeros_like(r) for r in res)
        gg[0][-1] = gy
        for q in range(n_aux):
            gg[1 + q][-1] = w[q]
        bst = []
        want = oo.odeint_split_rev(rev, tol, tol, math.inf, res, ts, (eps, p), gg, bst)
        return res, float(nfe_w), bst[0]["n_iters"], want

    res, nfe_w, bwd_w, want = oracle_run(cl["fwd"], cl["aug_dynamics"])
    resx, nfe_x, bwd_x, wantx = oracle_run(exact_arith(cl["fwd"]), exact_arith(cl["aug_dynamics"]))
    xg = x.to(cuda).requires_grad_(True)
    eg = eps.to(cuda).requires_grad_(True)
    pflat = spec.flatten_params(_to(p, cuda)).clone().requires_grad_(True)
    zc = z.to(cuda)
    fn = eno.odeint_sepaux if reg_type == "our" else eno.odeint_fin_sepaux
    outs, nfe = fn(spec.fwd, spec.aug_dynamics, (xg,) + (zc,) * n_aux, ts, eg, pflat, rtol=tol, atol=tol)
    loss = (outs[0][-1].reshape(gy.shape) * gy.to(cuda)).sum() + sum(w[q] * outs[1 + q][-1].sum() for q in range(n_aux))
    loss.backward()
    got_bwd = eno.last_stats["bwd"][0]["n_iters"]
    print(f"C4 split solve {reg_type}: nfe cuda {float(nfe):.0f} / fp32 oracle {nfe_w:.0f} / exact {nfe_x:.0f}; adjoint iterations "
          f"cuda {got_bwd} / {bwd_w} / {bwd_x}")
    assert_count(float(nfe), nfe_w, nfe_x, "nfe")
    assert_count(got_bwd, bwd_w, bwd_x, "adjoint iterations")
    ref_res, ref = (res, want) if (float(nfe), got_bwd) == (nfe_w, bwd_w) else (resx, wantx)
    assert rel_err(outs[0][-1].reshape(-1), ref_res[0][-1].reshape(-1)) < 1e-5
    assert rel_err(xg.grad.reshape(-1), ref[0][0].reshape(-1)) < 3e-5
    assert rel_err(eg.grad.reshape(-1), ref[2].reshape(-1)) < 3e-5
    assert rel_err(pflat.grad, spec.flatten_params(ref[3])) < 3e-5


def test_train_step_gradients_vs_oracle(cuda):
    """One update() of ffjord_mnist.py (pre-flow, split solve, bits/dim + lam mean(r), adjoint) on 12 x 12 images against
    the oracle's jax.grad restatement: step counts in the noise band, parameter gradients 3e-5."""
    from easy_neural_ode_b200 import ffjord_mnist_step as fs
    from oracle import ffjord_mnist_oracle as fm
    B, image, hidden, tol, lam = 3, (12, 12), 32, 1e-6, 3e-4
    p, _, eps = _case(B, image, hidden, seed=13, scale=1.3)
    images = torch.rand((B,) + image + (1,), generator=torch.Generator().manual_seed(3))
    model = fs.ImageFfjord(image, hidden, reg="r2", rtol=tol, atol=tol, device=cuda, lam=lam)
    model.params = model.spec.flatten_params(_to(p, cuda)).clone()
    out = model.loss_and_grads(images.to(cuda), eps.to(cuda))
    want = fm.train_step_grads(p, images, eps, lam=lam, reg="r2", rtol=tol, atol=tol, image=image + (1,))
    assert float(out["nfe"]) == float(want["nfe"])
    assert out["bwd_stats"][0]["n_iters"] == want["bwd_stats"][0]["n_iters"]
    assert rel_err(out["z"].reshape(-1), want["z"].reshape(-1)) < 1e-5
    assert abs(float(out["loss"]) - float(want["loss"])) < 1e-5 * abs(float(want["loss"]))
    assert rel_err(out["grads"], model.spec.flatten_params(want["grads"])) < 3e-5
