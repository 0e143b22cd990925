"""Host-side logic that needs no GPU: the dynamics-spec objects, parameter layout (must equal the
reference's ravel_pytree order), argument validation of the odeint surface, controller arithmetic
of the oracle on edge cases the reference's loop has (zero error, NaN)."""
import math

import pytest
import torch

import easy_neural_ode_b200 as eno
from oracle import dynamics_oracle as do, ode_oracle as oo
from oracle.jax_shim import ravel_pytree


def test_param_flattening_equals_ravel_pytree_order():
    spec = eno.MLPDynamics(16, 8, reg="r3")
    p = spec.init_params(seed=3)
    flat = spec.flatten_params(p)
    want, unravel = ravel_pytree(p)          # dict leaves in sorted-key order, like jax.flatten_util
    assert torch.equal(flat, want)
    back = spec.unflatten_params(flat)
    for k in p:
        for kk in p[k]:
            assert torch.equal(back[k][kk], p[k][kk])
    assert flat.numel() == spec.n_params


def test_init_matches_oracle_init():
    spec = eno.MLPDynamics(16, 8)
    a, b = spec.init_params(seed=5, scale=2.0), do.init_mlp_params(16, 8, seed=5, scale=2.0)
    for k in a:
        for kk in a[k]:
            assert torch.equal(a[k][kk], b[k][kk])


def test_spec_validation_and_flags():
    with pytest.raises(ValueError):
        eno.MLPDynamics(15, 8)
    with pytest.raises(ValueError):
        eno.MLPDynamics(16, 8, reg="r7")
    with pytest.raises(NotImplementedError):
        eno.MLPDynamics(16, 8, reg="r4")
    assert eno.MLPDynamics(16, 8, reg="none").reg_order == 0
    assert eno.MLPDynamics(16, 8, reg="r2").reg_order == 2
    assert eno.MLPDynamics(16, 8, reg="r3").reg_order == 3


def test_closures_cannot_be_called_and_callables_are_rejected():
    spec = eno.MLPDynamics(16, 8)
    p = spec.init_params()
    with pytest.raises(RuntimeError):
        spec.aug_dynamics((torch.zeros(2, 16), torch.zeros(2)), 0.0, None, p)
    with pytest.raises(TypeError):
        eno.odeint(lambda y, t, p: y, torch.zeros(2, 16), torch.tensor([0., 1.]), p)
    with pytest.raises(TypeError):
        eno.odeint_aux_one(spec.fwd, spec.fwd, (torch.zeros(2, 16), torch.zeros(2)), torch.tensor([0., 1.]), None, p)
    with pytest.raises(TypeError):
        eno.odeint(spec.dynamics_wrap, torch.zeros(2, 16), torch.tensor([0., 1.]))      # params missing


def test_controller_edge_cases():
    one = torch.tensor(0.1)
    assert float(oo.next_dt(one, torch.tensor(0.0))) == pytest.approx(1.0)              # ratio == 0 -> x10
    assert float(oo.next_dt(one, torch.tensor(1e-30))) == pytest.approx(1.0)            # growth clipped at 10
    assert float(oo.next_dt(one, torch.tensor(1.0))) == pytest.approx(0.1 * 0.9)        # accepted, yet shrinks
    assert float(oo.next_dt(one, torch.tensor(1e12))) == pytest.approx(0.02)            # shrink clipped at 5
    assert math.isnan(float(oo.next_dt(one, torch.tensor(float("nan")))))


def test_empty_time_grid_returns_initial_state_only():
    f = lambda y, t: -y
    ys, nfe, st = oo.solve(f, torch.ones(3), torch.tensor([0.5]), 1e-5, 1e-5)
    assert ys.shape == (1, 3) and nfe == 2.0 and st["n_iters"] == 0


def test_sepaux_surface_rejects_mismatched_closures():
    """odeint_sepaux mirrors lib/ode.py:678-697: (fwd_func, rev_func, (y, fro, kin), t, eps, params); the rev_func
    must be the 'fin' aug_dynamics of the same spec, and the usage errors surface before any device work."""
    import easy_neural_ode_b200 as eno
    our = eno.MLPDynamics(16, 8, reg="r3")
    fin = eno.MLPDynamics(16, 8, reg_type="fin")
    assert fin.aug_dynamics.kind == "fin" and our.aug_dynamics.kind == "aug"
    y0 = (torch.zeros(2, 16), torch.zeros(2), torch.zeros(2))
    ts = torch.tensor([0., 1.])
    eps = torch.ones(2, 16)
    p = fin.init_params()
    with pytest.raises(TypeError):
        eno.odeint_sepaux(our.fwd, our.aug_dynamics, y0, ts, eps, p)          # 'our' closure on the sepaux surface
    with pytest.raises(TypeError):
        eno.odeint_sepaux(fin.fwd, our.aug_dynamics, y0, ts, eps, p)          # closures of different specs
    with pytest.raises(TypeError):
        eno.odeint_sepaux(fin.fwd, fin.aug_dynamics, y0[:2], ts, eps, p)      # wrong state arity
    with pytest.raises(TypeError):
        eno.odeint_aux_one(fin.fwd, fin.aug_dynamics, y0[:2], ts, eps, p)     # 'fin' closure on the aux_one surface
    with pytest.raises(ValueError):
        eno.MLPDynamics(16, 8, reg_type="other")
