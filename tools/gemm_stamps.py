"""Development: where one tcgen05 GEMM launch of the C2 adjoint right-hand side spends its cycles (needs libeno_phases.so:
python easy-neural-ode_b200/build.py --phases).  Prints, for CTA (0,0,0) of the last launches, clock deltas between
entry / set-up done / first stage landed / last MMA issued / accumulators complete / epilogue done / exit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ENO_LIB"] = os.path.join(ROOT, "easy-neural-ode_b200", "libeno_phases.so")
import ctypes as C, torch
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import _cabi
dev = torch.device("cuda:0")
B, D, H = 1000, 43, 860
spec = eno.ConcatSquashDynamics(D, H, reg="r2")
p = spec.flatten_params(spec.init_params(seed=0, device=dev))
g = torch.Generator().manual_seed(0)
x, eps, yb = (torch.randn((B, D), generator=g).to(dev) for _ in range(3))
ab = torch.rand((2, B), generator=g).to(dev)
for _ in range(4):
    eno.rhs_closure(spec.aug_dynamics, 2, x, 0.3, eps, p, y_bar=yb, aux_bar=ab)
L = _cabi.lib()
L.eno_debug_gemm_stamps.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_uint)]
st = (C.c_longlong * 512)(); n = C.c_uint()
L.eno_debug_gemm_stamps(st, C.byref(n))
print("launches so far", n.value)
names = ["setup", "1st stage", "mainloop", "mma drain", "epilogue", "exit"]
print("  id  BN  kb grid      " + " ".join(f"{x:>10s}" for x in names) + "      total")
for k in range(min(n.value, 64)):
    i = (n.value - min(n.value, 64) + k) % 64
    s = st[8 * i: 8 * i + 8]
    meta = s[7]
    bn, kb, gx, gy = meta >> 40, (meta >> 24) & 0xFFFF, (meta >> 12) & 0xFFF, meta & 0xFFF
    d = [s[j + 1] - s[j] for j in range(6)]
    print(f"{k:4d} {bn:3d} {kb:3d} {gx:3d}x{gy:<3d}   " + " ".join(f"{x:10d}" for x in d) + f" {s[6] - s[0]:10d}")
