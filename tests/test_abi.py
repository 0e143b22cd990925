"""CPU-side checks of the drop-in boundary: libeno.so loads, exports every symbol include/eno.h
declares, agrees with the host binding on struct layouts and constants, and fails loudly (no CPU
fallback) when there is no sm_100a device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import _cabi
from oracle import tableaux as otab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "eno.h")).read()


def test_library_is_built_in_tree():
    assert os.path.exists(_cabi.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(_cabi.LIB_PATH).endswith("easy-neural-ode_b200")


def test_every_declared_symbol_is_exported():
    declared = set(re.findall(r"^(?:int32_t|int64_t|size_t|void|const char\*)\s+(eno_\w+)\(", HEADER, flags=re.M))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    L = _cabi.lib()
    for sym in declared:
        assert getattr(L, sym) is not None
    assert L.eno_abi_version() == int(re.search(r"#define ENO_ABI_VERSION (\d+)", HEADER).group(1))


def test_constants_match_header_and_oracle_ids():
    for name, idx in _cabi.TABLEAUX.items():
        m = re.search(rf"#define ENO_TABLEAU_{name.upper()} (\d+)", HEADER)
        assert m and int(m.group(1)) == idx
        assert otab.IDS[name] == idx
    assert int(re.search(r"#define ENO_NUM_TABLEAUX (\d+)", HEADER).group(1)) == len(_cabi.TABLEAUX)
    for k, v in dict(ENO_OK=0, ENO_MXSTEP=1, ENO_DT_UNDERFLOW=2, ENO_NONFINITE=3).items():
        assert int(re.search(rf"#define {k} (\d+)", HEADER).group(1)) == v


def test_struct_layouts():
    assert C.sizeof(_cabi.Desc) == 9 * 4
    assert C.sizeof(_cabi.Stats) == 8 * 4
    body = re.search(r"typedef struct eno_dynamics_desc \{(.*?)\} eno_dynamics_desc;", HEADER, re.S).group(1)
    assert len(re.findall(r"^\s*int32_t\s+\w+;", body, flags=re.M)) == len(_cabi.Desc._fields_)
    body = re.search(r"typedef struct eno_stats \{(.*?)\} eno_stats;", HEADER, re.S).group(1)
    assert len(re.findall(r"^\s*(?:int32_t|float)\s+\w+;", body, flags=re.M)) == len(_cabi.Stats._fields_)


def test_sizes_reported_by_the_library():
    L = _cabi.lib()
    spec = eno.MLPDynamics(784, 100, reg="r3")
    d = spec.desc(_cabi.AUG_REG, True)
    assert L.eno_param_count(C.byref(d)) == 158568 == spec.n_params          # P
    assert L.eno_adjoint_state_size(C.byref(d), 100) == 393969                # N_adj
    d0 = spec.desc(_cabi.AUG_NONE, False)
    assert L.eno_adjoint_state_size(C.byref(d0), 100) == 2 * 78400 + 1 + 158568
    assert L.eno_workspace_bytes(C.byref(d), 100, 2, 0, 0) > 8 * 78400 * 4
    assert L.eno_workspace_bytes(C.byref(d), 100, 2, 0, 1) > 0
    # 'fin' arm: two scalar states per sample and a live eps_bar block -> larger state, larger workspace
    dfin = spec.desc(_cabi.AUG_FIN, True, reg_order=0)
    assert L.eno_adjoint_state_size(C.byref(dfin), 100) == 2 * (78400 + 2 * 100) + 1 + 78400 + 158568
    assert L.eno_workspace_bytes(C.byref(dfin), 100, 2, 0, 1) > L.eno_workspace_bytes(C.byref(d), 100, 2, 0, 1)
    bad = _cabi.Desc(0, 783, 100, 3, 1, 1)
    assert L.eno_workspace_bytes(C.byref(bad), 100, 2, 0, 0) == 0              # D % 4 != 0 is unsupported


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    L = _cabi.lib()
    h = C.c_void_p()
    rc = L.eno_create(C.byref(h), 0)
    assert rc < 0 and b"CUDA" in L.eno_last_error_string(None)
    spec = eno.MLPDynamics(16, 8)
    p = spec.init_params()
    with pytest.raises(_cabi.EnoLibraryError):
        eno.odeint(spec.dynamics_wrap, torch.zeros(2, 16), torch.tensor([0., 1.]), p)


def test_null_arguments_are_usage_errors():
    L = _cabi.lib()
    assert L.eno_odeint_fwd(None, None, 0, 1, None, None, 2, None, None, 1e-5, 1e-5, 0, None, 0, None, None, None,
                            None, 0, None) == -1
    assert L.eno_odeint_bwd(None, None, 1, None, None, 2, None, None, None, 1e-5, 1e-5, 0, None, 0, None, None, None,
                            None, None, None, 0, None) == -1
    assert L.eno_profile_enable(None, 1) == -1
    assert L.eno_odeint_bwd_sepaux(None, None, 1, None, None, 2, None, None, None, None, 1e-5, 1e-5, 0, None, 0, None, None,
                                   None, None, None, None, None, 0, None) == -1
    assert L.eno_odeint_fwd_vmap(None, None, 1, None, None, 2, None, 1e-5, 1e-5, 0, None, None, None, None, None) == -1
    assert L.eno_mnist_head(None, 1, 4, 2, None, None, None, None, 1.0, None, None, None, None, None, None) == -1
    assert L.eno_momentum_update(None, 1, None, None, None, None, 0.1, 0.9, 0.0, None, None, None, None) == -1
