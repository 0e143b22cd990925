"""ctypes binding of include/eno.h (libeno.so).

There is no CPU or PyTorch fallback: if the library is missing, fails to load or
reports that the device is not sm_100a, every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ENO_LIB") or os.path.join(_HERE, "libeno.so")   # ENO_LIB: development variants only

ENO_OK, ENO_MXSTEP, ENO_DT_UNDERFLOW, ENO_NONFINITE = 0, 1, 2, 3
ARCH_MLP_SIGMOID, ARCH_CONCATSQUASH, ARCH_TANH_STACK, ARCH_CONCATCONV = 0, 1, 2, 3
F32, F64 = 0, 1
AUG_NONE, AUG_REG, AUG_ALL, AUG_FIN, AUG_FFJORD = 0, 1, 2, 3, 4
TABLEAUX = {"dopri5": 0, "heun": 1, "fehlberg": 2, "bosh": 3, "rk_fehlberg": 4, "cash_karp": 5,
            "owrenzen": 6, "owrenzen5": 7, "tanyam": 8}

# every symbol include/eno.h declares (checked by tests/test_abi.py)
SYMBOLS = ["eno_abi_version", "eno_last_error_string", "eno_create", "eno_destroy", "eno_param_count",
           "eno_adjoint_state_size", "eno_workspace_bytes", "eno_odeint_fwd", "eno_odeint_fwd_vmap", "eno_odeint_bwd", "eno_odeint_bwd_sepaux", "eno_odeint_grid_fwd", "eno_odeint_grid_bwd", "eno_rhs",
           "eno_mnist_head", "eno_momentum_update", "eno_profile_enable", "eno_profile_read", "eno_ffma_peak", "eno_comm_unique_id", "eno_comm_init",
           "eno_comm_destroy", "eno_comm_set_shards", "eno_set_deferred", "eno_read_stats", "eno_gemm_tf32x3", "eno_gemm_tf32x3_mn", "eno_adam_update", "eno_profile_read_gemm", "eno_peer_export", "eno_peer_attach",
           "eno_traj_param_count", "eno_traj_workspace_bytes", "eno_traj_fwd", "eno_traj_bwd", "eno_traj_rhs"]


class Desc(C.Structure):
    _fields_ = [("arch", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("reg_order", C.c_int32),
                ("aug", C.c_int32), ("eps_in_args", C.c_int32), ("n_hidden", C.c_int32), ("img_h", C.c_int32),
                ("img_w", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("nfe", C.c_float), ("n_iters", C.c_int32), ("n_accepted", C.c_int32), ("status", C.c_int32),
                ("last_dt", C.c_float), ("last_t", C.c_float), ("n_launches", C.c_int32), ("n_rhs", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


class EnoLibraryError(RuntimeError):
    pass


def lib():
    """Loads libeno.so once; raises EnoLibraryError when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EnoLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the ODE path has no CPU / PyTorch fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
    L.eno_abi_version.restype = i32
    L.eno_last_error_string.restype = C.c_char_p
    L.eno_last_error_string.argtypes = [vp]
    L.eno_create.restype = i32
    L.eno_create.argtypes = [C.POINTER(vp), i32]
    L.eno_destroy.restype = None
    L.eno_destroy.argtypes = [vp]
    L.eno_param_count.restype = i64
    L.eno_param_count.argtypes = [C.POINTER(Desc)]
    L.eno_adjoint_state_size.restype = i64
    L.eno_adjoint_state_size.argtypes = [C.POINTER(Desc), i32]
    L.eno_workspace_bytes.restype = sz
    L.eno_workspace_bytes.argtypes = [C.POINTER(Desc), i32, i32, i32, i32]
    L.eno_odeint_fwd.restype = i32
    L.eno_odeint_fwd.argtypes = [vp, C.POINTER(Desc), i32, i32, vp, vp, i32, vp, vp, f32, f32, i32, vp, sz, vp, vp,
                                 C.POINTER(Stats), vp, i32, vp]
    L.eno_odeint_fwd_vmap.restype = i32
    L.eno_odeint_fwd_vmap.argtypes = [vp, C.POINTER(Desc), i32, vp, vp, i32, vp, f32, f32, i32, vp, vp, vp, vp, vp]
    L.eno_odeint_bwd.restype = i32
    L.eno_odeint_bwd.argtypes = [vp, C.POINTER(Desc), i32, vp, vp, i32, vp, vp, vp, f32, f32, i32, vp, sz, vp, vp, vp,
                                 vp, C.POINTER(Stats), vp, i32, vp]
    L.eno_odeint_bwd_sepaux.restype = i32
    L.eno_odeint_bwd_sepaux.argtypes = [vp, C.POINTER(Desc), i32, vp, vp, i32, vp, vp, vp, vp, f32, f32, i32, vp, sz, vp, vp,
                                        vp, vp, vp, C.POINTER(Stats), vp, i32, vp]
    L.eno_odeint_grid_fwd.restype = i32
    L.eno_odeint_grid_fwd.argtypes = [vp, C.POINTER(Desc), i32, vp, vp, i32, vp, vp, f32, vp, sz, vp, vp, C.POINTER(Stats), vp]
    L.eno_odeint_grid_bwd.restype = i32
    L.eno_odeint_grid_bwd.argtypes = [vp, C.POINTER(Desc), i32, vp, vp, i32, vp, vp, vp, vp, f32, vp, sz, vp, vp, vp, vp, vp,
                                      C.POINTER(Stats), vp]
    L.eno_rhs.restype = i32
    L.eno_rhs.argtypes = [vp, C.POINTER(Desc), i32, i32, vp, f32, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp, vp, vp, vp]
    L.eno_mnist_head.restype = i32
    L.eno_mnist_head.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp]
    L.eno_momentum_update.restype = i32
    L.eno_momentum_update.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), f32, f32, f32, vp, vp, vp, vp]
    L.eno_gemm_tf32x3_mn.restype = i32
    L.eno_gemm_tf32x3_mn.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32, vp, i32, i32, i32, vp]
    L.eno_comm_set_shards.restype = i32
    L.eno_comm_set_shards.argtypes = [vp, i32, C.c_int64]
    L.eno_profile_enable.restype = i32
    L.eno_profile_enable.argtypes = [vp, i32]
    L.eno_profile_read.restype = i32
    L.eno_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.eno_ffma_peak.restype = i32
    L.eno_ffma_peak.argtypes = [vp, C.POINTER(C.c_double), vp]
    L.eno_comm_unique_id.restype = i32
    L.eno_comm_unique_id.argtypes = [vp, vp, C.c_char_p]
    L.eno_comm_init.restype = i32
    L.eno_comm_init.argtypes = [vp, vp, i32, i32, C.c_char_p]
    L.eno_comm_destroy.restype = i32
    L.eno_comm_destroy.argtypes = [vp]
    L.eno_set_deferred.restype = i32
    L.eno_set_deferred.argtypes = [vp, i32, i32, i32]
    L.eno_read_stats.restype = i32
    L.eno_read_stats.argtypes = [vp, vp, C.POINTER(Stats), vp]
    L.eno_peer_export.restype = i32
    L.eno_peer_export.argtypes = [vp, C.c_uint64, vp]
    L.eno_peer_attach.restype = i32
    L.eno_peer_attach.argtypes = [vp, vp]
    L.eno_profile_read_gemm.restype = i32
    L.eno_profile_read_gemm.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.eno_adam_update.restype = i32
    L.eno_adam_update.argtypes = [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp]
    L.eno_gemm_tf32x3.restype = i32
    L.eno_gemm_tf32x3.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32, vp, i32, i32, i32, vp]
    f64 = C.c_double
    L.eno_traj_param_count.restype = i64
    L.eno_traj_param_count.argtypes = [C.POINTER(Desc)]
    L.eno_traj_workspace_bytes.restype = sz
    L.eno_traj_workspace_bytes.argtypes = [C.POINTER(Desc), i32, i32]
    L.eno_traj_fwd.restype = i32
    L.eno_traj_fwd.argtypes = [vp, C.POINTER(Desc), i32, i32, vp, vp, i32, vp, f64, f64, i32, vp, vp, vp, vp, vp, vp]
    L.eno_traj_bwd.restype = i32
    L.eno_traj_bwd.argtypes = [vp, C.POINTER(Desc), i32, i32, vp, vp, vp, i32, vp, vp, vp, f64, f64, i32, vp, sz, vp, vp, vp,
                               vp, vp, vp, vp, vp]
    L.eno_traj_rhs.restype = i32
    L.eno_traj_rhs.argtypes = [vp, C.POINTER(Desc), i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


_ctxs = {}


def ctx(device_index):
    """One eno_ctx per CUDA device, created on first use."""
    if device_index not in _ctxs:
        L = lib()
        h = C.c_void_p()
        rc = L.eno_create(C.byref(h), int(device_index))
        if rc != 0:
            raise EnoLibraryError(f"eno_create failed ({rc}): {L.eno_last_error_string(None).decode()}")
        _ctxs[device_index] = h
    return _ctxs[device_index]


def check(rc, handle, what):
    if rc < 0:
        msg = lib().eno_last_error_string(handle).decode()
        raise EnoLibraryError(f"{what} failed ({rc}): {msg}")
    return rc
