"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitize_case.py
ENO_PATH=stream runs the streamed formulation instead.  Cluster-resident MLP kernels (DSMEM all-reduce after cluster.sync, aliased dead buffers, ticket-elected finalisers), the
'fin' adjoint, the per-example solve, the streamed path, and the tcgen05 GEMM pipeline of the ConcatSquash dynamics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, D, H, tol = 8, 64, 16, 1e-5
ts = torch.tensor([0., 1.], device=dev)
x0 = torch.rand((B, D), generator=g).to(dev)
eps = (torch.randint(0, 2, (B, D), generator=g).float() * 2 - 1).to(dev)
gy = torch.randn((B, D), generator=g).to(dev)


def run_mlp(reg, reg_type):
    spec = eno.MLPDynamics(D, H, reg=reg, reg_type=reg_type)
    p = spec.flatten_params(spec.init_params(seed=0, device=dev, scale=2.0)).clone().requires_grad_(True)
    xg = x0.clone().requires_grad_(True)
    z = torch.zeros(B, device=dev)
    if reg_type == "fin":
        (ys, fro, kin), nfe = eno.odeint_sepaux(spec.fwd, spec.aug_dynamics, (xg, z, z), ts, eps, p, rtol=tol, atol=tol)
        loss = (ys[-1] * gy).sum() + 0.01 * fro[-1].sum() + 0.02 * kin[-1].sum()
    else:
        (ys, rs), nfe = eno.odeint_aux_one(spec.fwd, spec.aug_dynamics, (xg, z), ts, eps, p, rtol=tol, atol=tol)
        loss = (ys[-1] * gy).sum() + 0.01 * rs[-1].sum()
    loss.backward()
    torch.cuda.synchronize()
    print(f"mlp reg={reg} reg_type={reg_type} path={os.environ.get('ENO_PATH', 'cluster')}: nfe {float(nfe)} bwd iters "
          f"{eno.last_stats['bwd'][0]['n_iters']}", flush=True)
    return spec, p


spec, p = run_mlp("r3", "our")
if os.environ.get("ENO_PATH") == "stream":     # the streamed formulation covers plain / R2 / R3 only
    print("sanitize_case finished (streamed path)")
    sys.exit(0)
run_mlp("none", "fin")
ys, nfe = eno.odeint_vmap(spec.dynamics_wrap, x0, ts, p.detach(), rtol=tol, atol=tol)
torch.cuda.synchronize()
print("vmap nfe", nfe.tolist(), flush=True)
# evaluation path (augmented forward)
z = torch.zeros(B, device=dev)
out, nfe = eno.odeint(spec.all_aug_dynamics, (x0, z, z, z), ts, eps, p.detach(), rtol=tol, atol=tol)
torch.cuda.synchronize()
print("all_aug forward nfe", float(nfe), flush=True)
# ConcatSquash / tcgen05 pipeline
Bc, Dc, Hc = 40, 6, 24
cs = eno.ConcatSquashDynamics(Dc, Hc, reg="r2")
pc = cs.flatten_params(cs.init_params(seed=0, device=dev, scale=1.5)).clone().requires_grad_(True)
xc = torch.randn((Bc, Dc), generator=g).to(dev).requires_grad_(True)
ec = torch.randn((Bc, Dc), generator=g).to(dev)
zc = torch.zeros(Bc, 1, device=dev)
(ys, dl, rs), nfe = eno.odeint_sepaux(cs.fwd, cs.aug_dynamics, (xc, zc, zc), ts, ec, pc, rtol=tol, atol=tol)
(ys[-1].sum() + dl[-1].sum() / Bc + 0.01 * rs[-1].sum() / Bc).backward()
torch.cuda.synchronize()
print("concatsquash nfe", float(nfe), "bwd iters", eno.last_stats["bwd"][0]["n_iters"], flush=True)
print("sanitize_case finished")
