"""GPU parity of the fixed-grid RK4 family (odeint_grid, odeint_grid_sepaux_one, odeint_grid_sepaux; lib/ode.py:561-634,
864-946, 1531-1593) on the flat-state engine, through the C-ABI (eno_odeint_grid_fwd / eno_odeint_grid_bwd):
against tests/golden/css_grid.npz (the unmodified reference solver) and against the oracle at a mid size and for the
ConcatConv2D closures.  Step counts (nfe = 4 per step) exact, values and gradients rtol 1e-5 (north_star)."""
import pytest
import torch

from tests.util import load, rel_err

pytestmark = pytest.mark.gpu


def _to(p, dev, dtype=torch.float32):
    return {k: {kk: vv.to(device=dev, dtype=dtype) for kk, vv in v.items()} for k, v in p.items()}


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("reg_type", ["our", "fin"])
def test_grid_split_solves_golden(cuda, tag, reg_type):
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo
    g = load("css_grid.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    p = fo.init_css_params(D, H, seed=int(g["seed"]), scale=float(g["scale"]))
    step = float(g[f"{tag}_step"])
    key = tag if reg_type == "our" else f"{tag}_fin"
    n_aux = 2 if reg_type == "our" else 3
    spec = eno.ConcatSquashDynamics(D, H, reg="r2" if reg_type == "our" else "none", reg_type=reg_type)
    x0 = torch.as_tensor(g["x0"], device=cuda).requires_grad_(True)
    eps = torch.as_tensor(g["eps"], device=cuda).requires_grad_(True)
    pflat = spec.flatten_params(_to(p, cuda)).clone().requires_grad_(True)
    z = torch.zeros(B, 1, device=cuda)
    fn = eno.odeint_grid_sepaux_one if reg_type == "our" else eno.odeint_grid_sepaux
    outs, nfe = fn(spec.fwd, spec.aug_dynamics, (x0,) + (z,) * n_aux, torch.tensor([0., 1.]), eps, pflat, step_size=step)
    assert float(nfe) == float(g[f"{key}_nfe"])
    assert eno.last_stats["fwd"]["n_iters"] == 4 and eno.last_stats["fwd"]["n_rhs"] == 17
    assert all(tuple(a.shape) == (2, B, 1) and float(a.detach().abs().max()) == 0.0 for a in outs[1:])
    assert rel_err(outs[0].detach(), g[f"{key}_y"]) < 1e-5
    w = [1.0 / B, 0.01 / B, 0.02 / B]
    loss = (outs[0][-1] * torch.as_tensor(g["gy"], device=cuda)).sum() + sum(w[q] * outs[1 + q][-1].sum() for q in range(n_aux))
    loss.backward()
    assert eno.last_stats["bwd"][0]["n_iters"] == 4
    sort_flat = lambda tree: torch.cat([v.reshape(-1) for n in spec.layer_names for _, v in sorted(tree[n].items())])
    assert rel_err(x0.grad, g[f"{key}_x0_bar"]) < 1e-5
    assert rel_err(eps.grad, g[f"{key}_eps_bar"]) < 1e-5
    assert rel_err(sort_flat(spec.unflatten_params(pflat.grad.cpu())), g[f"{key}_params_bar"]) < 1e-5


@pytest.mark.parametrize("tag", ["a", "b"])
def test_plain_grid_solve_golden(cuda, tag):
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo
    g = load("css_grid.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    p = fo.init_css_params(D, H, seed=int(g["seed"]), scale=float(g["scale"]))
    spec = eno.ConcatSquashDynamics(D, H)
    ys, nfe = eno.odeint_grid(spec.fwd, torch.as_tensor(g["x0"], device=cuda), torch.tensor([0., 1.]),
                              torch.as_tensor(g["eps"], device=cuda), spec.flatten_params(_to(p, cuda)), step_size=float(g[f"{tag}_step"]))
    assert float(nfe) == float(g[f"{tag}_plain_nfe"]) and tuple(ys.shape) == (2, B, D)
    assert rel_err(ys, g[f"{tag}_plain_y"]) < 1e-5
    assert torch.equal(ys[0].cpu(), torch.as_tensor(g["x0"]))


def test_grid_mid_size_vs_oracle(cuda):
    """ConcatSquash 43 -> 128 -> 128 -> 43, batch 64, ten steps of 0.1 (+ the clipped-step case 0.15): nfe and every
    gradient against the oracle."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_oracle as fo, ode_oracle as oo
    B, D, H = 64, 43, 128
    p = fo.init_css_params(D, H, n_hidden_layers=2, seed=3, scale=1.3)
    gen = torch.Generator().manual_seed(9)
    x, eps, gy = (torch.randn((B, D), generator=gen) for _ in range(3))
    spec = eno.ConcatSquashDynamics(D, H, reg="r2")
    cl = fo.make_ffjord_closures("r2", "our")
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    for step in (0.1, 0.15):
        res, nfe_w = oo.odeint_grid_split_fwd(cl["fwd"], (x, z, z), ts, eps, p, n_aux=2, step_size=step)
        gg = tuple(torch.zeros_like(r) for r in res)
        gg[0][-1], gg[1][-1], gg[2][-1] = gy / B, 1.0 / B, 0.01 / B
        want = oo.odeint_grid_split_rev(cl["aug_dynamics"], step, res, ts, (eps, p), gg)
        xg = x.to(cuda).requires_grad_(True)
        eg = eps.to(cuda).requires_grad_(True)
        pflat = spec.flatten_params(_to(p, cuda)).clone().requires_grad_(True)
        zc = z.to(cuda)
        (ys, dl, rs), nfe = eno.odeint_grid_sepaux_one(spec.fwd, spec.aug_dynamics, (xg, zc, zc), ts, eg, pflat, step_size=step)
        assert float(nfe) == nfe_w
        ((ys[-1] * gy.to(cuda)).sum() / B + dl[-1].sum() / B + 0.01 * rs[-1].sum() / B).backward()
        assert rel_err(ys.detach(), res[0]) < 1e-5
        assert rel_err(xg.grad, want[0][0]) < 1e-5
        assert rel_err(eg.grad, want[2]) < 1e-5
        assert rel_err(pflat.grad, spec.flatten_params(want[3])) < 1e-5


def test_grid_conv_vs_oracle(cuda):
    """ffjord_mnist.py's fixed-grid call (ffjord_mnist.py:68-71) on 12 x 12 images, 32 channels, five steps of 0.2."""
    import easy_neural_ode_b200 as eno
    from oracle import ffjord_mnist_oracle as fm, ode_oracle as oo
    B, image, hidden, step = 3, (12, 12), 32, 0.2
    img = image + (1,)
    p = fm.init_conv_params(hidden=hidden, seed=11, scale=1.3, last_scale=1.0)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn((B,) + img, generator=gen)
    eps = torch.randint(0, 2, (B,) + img, generator=gen).float() * 2 - 1
    gy = torch.randn((B,) + img, generator=gen)
    spec = eno.ConcatConvDynamics(image, hidden, reg="r2")
    cl = fm.make_ffjord_mnist_closures("r2", "our", img)
    ts = torch.tensor([0., 1.])
    z = torch.zeros(B, 1)
    res, nfe_w = oo.odeint_grid_split_fwd(cl["fwd"], (x, z, z), ts, eps, p, n_aux=2, step_size=step)
    gg = tuple(torch.zeros_like(r) for r in res)
    gg[0][-1], gg[1][-1], gg[2][-1] = gy, 1.0 / B, 0.01 / B
    want = oo.odeint_grid_split_rev(cl["aug_dynamics"], step, res, ts, (eps, p), gg)
    xg = x.to(cuda).requires_grad_(True)
    eg = eps.to(cuda).requires_grad_(True)
    pflat = spec.flatten_params(_to(p, cuda)).clone().requires_grad_(True)
    zc = z.to(cuda)
    outs, nfe = eno.odeint_grid_sepaux_one(spec.fwd, spec.aug_dynamics, (xg, zc, zc), ts, eg, pflat, step_size=step)
    assert float(nfe) == nfe_w == 20.0
    ((outs[0][-1].reshape(gy.shape) * gy.to(cuda)).sum() + outs[1][-1].sum() / B + 0.01 * outs[2][-1].sum() / B).backward()
    assert rel_err(outs[0][-1].detach().reshape(-1), res[0][-1].reshape(-1)) < 1e-5
    assert rel_err(xg.grad.reshape(-1), want[0][0].reshape(-1)) < 2e-5
    assert rel_err(eg.grad.reshape(-1), want[2].reshape(-1)) < 2e-5
    assert rel_err(pflat.grad, spec.flatten_params(want[3])) < 2e-5


def test_grid_rejects_unsupported(cuda):
    import easy_neural_ode_b200 as eno
    mlp = eno.MLPDynamics(16, 8)
    with pytest.raises(NotImplementedError):
        eno.odeint_grid(mlp.dynamics_wrap, torch.zeros(2, 16, device=cuda), torch.tensor([0., 1.]), mlp.init_params(device=cuda), step_size=0.1)
    spec = eno.ConcatSquashDynamics(6, 12)
    p = spec.flatten_params(spec.init_params(device=cuda))
    with pytest.raises(Exception):   # three output times: the reference's loop only knows ts[-1] (lib/ode.py:884)
        eno.odeint_grid(spec.fwd, torch.zeros(2, 6, device=cuda), torch.tensor([0., 0.5, 1.]), torch.zeros(2, 6, device=cuda), p, step_size=0.1)
