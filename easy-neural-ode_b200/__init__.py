"""B200-native adaptive ODE-solve path of easy-neural-ode (odeint + adjoint + Taylor regulariser).

Import name: ``easy_neural_ode_b200`` (a one-file shim next to this directory maps it here,
because the directory name the project layout prescribes contains a hyphen).
"""
from . import _cabi
from .dynamics import MLPDynamics, ConcatSquashDynamics, ConcatConvDynamics, GenDynamics, EnoFunc, REGS
from . import latent
from .ode import (odeint, odeint_aux_one, odeint_sepaux, odeint_fin_sepaux, odeint_vmap, odeint_grid, odeint_grid_sepaux_one,
                  odeint_grid_sepaux, _dopri5_odeint, _heun_odeint, _fehlberg_odeint, _bosh_odeint,
                  _rk_fehlberg_odeint, _cash_karp_odeint, _owrenzen_odeint, _owrenzen5_odeint, _tanyam_odeint,
                  solve_with_trace, rhs, rhs_closure, last_stats, init_distributed, shutdown_distributed, shard_rows, set_shards, set_deferred,
                  read_deferred_stats)

__all__ = ["MLPDynamics", "ConcatSquashDynamics", "ConcatConvDynamics", "GenDynamics", "EnoFunc", "REGS", "odeint", "odeint_aux_one", "odeint_sepaux", "odeint_fin_sepaux", "odeint_vmap", "odeint_grid", "odeint_grid_sepaux_one", "odeint_grid_sepaux", "solve_with_trace", "rhs", "last_stats", "init_distributed", "shutdown_distributed", "shard_rows", "set_shards"]
