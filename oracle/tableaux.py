"""Butcher tableaux of the reference's explicit embedded RK pairs, as data.

TEST INFRASTRUCTURE (oracle).  Values restate /root/reference/lib/ode.py:
dopri5 108-120 (+ mid-point row 58-61), bosh 138-145, heun 163-168, fehlberg
186-192, rk_fehlberg 210-219, cash_karp 237-246, owrenzen 264-273, owrenzen5
291-302, tanyam 320-364.  All entries are Python doubles evaluated exactly as
the reference spells them (e.g. ``35/384 - 1951/21600``) and are rounded to the
working dtype only when a solve starts; the one exception is tanyam's ``c_sol``,
which the reference forms by adding two *already rounded* arrays
(lib/ode.py:350-359) - ``as_dtype`` reproduces that.

Per-solver constants: ``order`` feeds
optimal_step_size, ``init_order`` feeds initial_step_size, ``nfe_per_iter`` is
what the reference adds to its NFE counter per loop iteration.

The CUDA path does NOT import this file: it carries its own constexpr copy in
``easy-neural-ode_b200/csrc/tableaux.cuh`` so the oracle stays an independent
checker.
"""
import numpy as np

_T = {}

_T["dopri5"] = dict(
    alpha=[1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1., 0],
    beta=[
        [1 / 5, 0, 0, 0, 0, 0, 0],
        [3 / 40, 9 / 40, 0, 0, 0, 0, 0],
        [44 / 45, -56 / 15, 32 / 9, 0, 0, 0, 0],
        [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0, 0, 0],
        [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656, 0, 0],
        [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
    ],
    c_sol=[35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0],
    c_err=[35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085,
           125 / 192 - 451 / 720, -2187 / 6784 - -12231 / 42400,
           11 / 84 - 649 / 6300, -1. / 60.],
    c_mid=[6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
           -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
           -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2],
    order=5, init_order=4, nfe_per_iter=6, clip=False,
)

_T["heun"] = dict(
    alpha=[1., 0],
    beta=[[1 / 2, 0]],
    c_sol=[1 / 2, 1 / 2],
    c_err=[1 / 2 - 1, 1 / 2],
    order=2, init_order=1, nfe_per_iter=1, clip=True,
)

_T["fehlberg"] = dict(
    alpha=[1 / 2, 1, 0],
    beta=[[1 / 2, 0, 0], [1 / 256, 255 / 256, 0]],
    c_sol=[1 / 512, 255 / 256, 1 / 512],
    c_err=[1 / 512 - 1 / 256, 0., 1 / 512],
    order=2, init_order=1, nfe_per_iter=2, clip=True,
)

_T["bosh"] = dict(
    alpha=[1 / 2, 3 / 4, 1., 0],
    beta=[[1 / 2, 0, 0, 0], [0., 3 / 4, 0, 0], [2 / 9, 1 / 3, 4 / 9, 0]],
    c_sol=[2 / 9, 1 / 3, 4 / 9, 0.],
    c_err=[2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8],
    order=3, init_order=2, nfe_per_iter=3, clip=True,
)

_T["rk_fehlberg"] = dict(
    alpha=[1 / 4, 3 / 8, 12 / 13, 1, 1 / 2, 0],
    beta=[
        [1 / 4, 0, 0, 0, 0, 0],
        [3 / 32, 9 / 32, 0, 0, 0, 0],
        [1932 / 2197, -7200 / 2197, 7296 / 2197, 0, 0, 0],
        [439 / 216, -8, 3680 / 513, -845 / 4104, 0, 0],
        [-8 / 27, 2, -3544 / 2565, 1859 / 4104, -11 / 40, 0],
    ],
    c_sol=[16 / 135, 0, 6656 / 12825, 28561 / 56430, -9 / 50, 2 / 55],
    c_err=[16 / 135 - 25 / 216, 0, 6656 / 12825 - 1408 / 2565,
           28561 / 56430 - 2197 / 4104, -9 / 50 - - 1 / 5, 2 / 55],
    order=5, init_order=4, nfe_per_iter=5, clip=True,
)

_T["cash_karp"] = dict(
    alpha=[1 / 5, 3 / 10, 3 / 5, 1, 7 / 8, 0],
    beta=[
        [1 / 5, 0, 0, 0, 0, 0],
        [3 / 40, 9 / 40, 0, 0, 0, 0],
        [3 / 10, -9 / 10, 6 / 5, 0, 0, 0],
        [-11 / 54, 5 / 2, -70 / 27, 35 / 27, 0, 0],
        [1631 / 55296, 175 / 512, 575 / 13824, 44275 / 110592, 253 / 4096, 0],
    ],
    c_sol=[37 / 378, 0, 250 / 621, 125 / 594, 0, 512 / 1771],
    c_err=[37 / 378 - 2825 / 27648, 0, 250 / 621 - 18575 / 48384,
           125 / 594 - 13525 / 55296, -277 / 14336, 512 / 1771 - 1 / 4],
    order=5, init_order=4, nfe_per_iter=5, clip=True,
)

_T["owrenzen"] = dict(
    alpha=[1 / 6, 11 / 37, 11 / 17, 13 / 15, 1, 0],
    beta=[
        [1 / 6, 0, 0, 0, 0, 0],
        [44 / 1369, 363 / 1369, 0, 0, 0, 0],
        [3388 / 4913, -8349 / 4913, 8140 / 4913, 0, 0, 0],
        [-36764 / 408375, 767 / 1125, -32708 / 136125, 210392 / 408375, 0, 0],
        [1697 / 18876, 0, 50653 / 116160, 299693 / 1626240, 3375 / 11648, 0],
    ],
    c_sol=[1697 / 18876, 0, 50653 / 116160, 299693 / 1626240, 3375 / 11648, 0],
    c_err=[-1185 / 6292, 0, 4107 / 7744, -68493 / 108416, 3375 / 11648, 0],
    order=4, init_order=3, nfe_per_iter=5, clip=True,
)

_T["owrenzen5"] = dict(
    alpha=[1 / 6, 1 / 4, 1 / 2, 1 / 2, 9 / 14, 7 / 8, 1, 0],
    beta=[
        [1 / 6, 0, 0, 0, 0, 0, 0, 0],
        [1 / 16, 3 / 16, 0, 0, 0, 0, 0, 0],
        [1 / 4, -3 / 4, 1, 0, 0, 0, 0, 0],
        [-3 / 4, 15 / 4, -3, 1 / 2, 0, 0, 0, 0],
        [369 / 1372, -243 / 343, 297 / 343, 1485 / 9604, 297 / 4802, 0, 0, 0],
        [-133 / 4512, 1113 / 6016, 7945 / 16544, -12845 / 24064, -315 / 24064,
         156065 / 198528, 0, 0],
        [83 / 945, 0, 248 / 825, 41 / 180, 1 / 36, 2401 / 38610, 6016 / 20475, 0],
    ],
    c_sol=[-1 / 9 + 188 / 945, 0, 40 / 33 - 752 / 825, -7 / 4 + 89 / 45, -1 / 12 + 1 / 9,
           343 / 198 - 32242 / 19305, 6016 / 20475, 0],
    c_err=[188 / 945, 0, -752 / 825, 89 / 45, 1 / 9, -32242 / 19305, 6016 / 20475, 0],
    order=5, init_order=4, nfe_per_iter=7, clip=True,
)

_TY_BHAT = [0.05126014249744686, 0, 0, 0.27521638456212627, 0.33696650340710543,
            0.18986072244906577, 8.461099418514403, -130.15941672640542,
            121.84501355497527, 0]
_TY_ERR = [0.0002577249070696835, 0, 0, -0.0009229104845391819, 0.0031779013374194105,
           -0.01210458894174817, 2.706023959472591, -44.541445641266066,
           121.84501355497527, -80]
_T["tanyam"] = dict(
    alpha=[0.07816646510113846, 0.1172496976517077, 0.17587454647756157,
           0.4987401101913988, 0.772121690184484, 0.9911856696047768,
           0.9995019582097662, 1, 0, 0],
    beta=[
        [0.07816646510113846, 0, 0, 0, 0, 0, 0, 0, 0, 0],
        [0.029312424412926925, 0.08793727323878078, 0, 0, 0, 0, 0, 0, 0, 0],
        [0.04396863661939039, 0, 0.13190590985817116, 0, 0, 0, 0, 0, 0, 0],
        [0.7361834837738382, 0, -2.8337999624233303, 2.5963565888408913, 0, 0, 0, 0, 0, 0],
        [-12.062819391370866, 0, 48.20838100175243, -38.05863046463434,
         2.6851905444372632, 0, 0, 0, 0, 0],
        [105.21957276320198, 0, -417.92888626241256, 332.3155504499333,
         -19.827591183572938, 1.2125399024549859, 0, 0, 0, 0],
        [114.67755718631742, 0, -455.5612169896097, 362.24095553923144,
         -21.67190442182809, 1.3189132007137807, -0.0048025566150346555, 0, 0, 0],
        [115.21334870553768, 0, -457.69356568613233, 363.93688218862735,
         -21.776682078900294, 1.3250670887878468, -0.004518190986768983,
         -0.0005320269334859959, 0, 0],
        [115.18928245800194, 0, -457.598022271643, 363.8610256312148,
         -21.77212754027556, 1.3248804645074317, -0.0045057252106918315,
         -0.0005330165949429136, 0, 0],
    ],
    c_sol=None,  # formed in the working dtype, see as_dtype
    c_err=_TY_ERR,
    order=5, init_order=4, nfe_per_iter=9, clip=True,
)

NAMES = ["dopri5", "heun", "fehlberg", "bosh", "rk_fehlberg", "cash_karp",
         "owrenzen", "owrenzen5", "tanyam"]
# Integer ids; they mirror ENO_TABLEAU_* in include/eno.h (checked in tests).
IDS = {n: i for i, n in enumerate(NAMES)}
MAX_STAGES = 10


def stages(name):
    return len(_T[name]["alpha"])


def meta(name):
    t = _T[name]
    return dict(stages=stages(name), order=t["order"], init_order=t["init_order"],
                nfe_per_iter=t["nfe_per_iter"], clip=t["clip"])


def as_dtype(name, dtype=np.float32):
    """Tableau rounded to ``dtype`` the way the reference rounds it."""
    t = _T[name]
    s = stages(name)
    dt = np.dtype(dtype)
    out = dict(meta(name))
    out["alpha"] = np.asarray(t["alpha"], dtype=np.float64).astype(dt)
    beta = np.zeros((s - 1, s), dtype=np.float64)
    beta[:, :] = np.asarray(t["beta"], dtype=np.float64)
    out["beta"] = beta.astype(dt)
    if name == "tanyam":
        out["c_sol"] = np.asarray(_TY_BHAT, np.float64).astype(dt) + np.asarray(_TY_ERR, np.float64).astype(dt)
    else:
        out["c_sol"] = np.asarray(t["c_sol"], dtype=np.float64).astype(dt)
    out["c_err"] = np.asarray(t["c_err"], dtype=np.float64).astype(dt)
    if "c_mid" in t:
        out["c_mid"] = np.asarray(t["c_mid"], dtype=np.float64).astype(dt)
    return out
