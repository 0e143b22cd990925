"""Timing of the per-example (--vmap) NFE path: eno.odeint_vmap on B200 vs the oracle solving every sample alone
on the host.  Usage: python tools/vmap_time.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno
from oracle import dynamics_oracle as do, ode_oracle as oo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
D, H = 784, 100
dev = torch.device("cuda:0")
spec = eno.MLPDynamics(D, H)
p_cpu = do.init_mlp_params(D, H, seed=0)
params = {k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in p_cpu.items()}
gen = torch.Generator().manual_seed(0)
x0 = (torch.randint(0, 256, (B, D), generator=gen).float() / 255.)
ts = torch.tensor([0., 1.])
xd = x0.to(dev)
for tol in (1.4e-8, 1e-5):
    ys, nfe = eno.odeint_vmap(spec.dynamics_wrap, xd, ts, params, rtol=tol, atol=tol)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ys, nfe = eno.odeint_vmap(spec.dynamics_wrap, xd, ts, params, rtol=tol, atol=tol)
    e1.record(); torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 10
    _, nfe_b = eno.odeint(spec.dynamics_wrap, xd, ts, params, rtol=tol, atol=tol)
    n_cpu = min(B, 8)
    t0 = time.perf_counter()
    want = [float(oo.odeint(do.mlp_dynamics_xtp, x0[b:b + 1], ts, p_cpu, rtol=tol, atol=tol)[1]) for b in range(n_cpu)]
    cpu_ms = (time.perf_counter() - t0) * 1e3 / n_cpu * B
    print(f"tol {tol:g}: B={B} per-example solves in {gpu_ms:.3f} ms (one launch); mean NFE {float(nfe.mean()):.2f} "
          f"min {float(nfe.min()):.0f} max {float(nfe.max()):.0f}; batch solve NFE {float(nfe_b):.0f}; "
          f"first {n_cpu} samples match oracle NFE: {nfe[:n_cpu].cpu().tolist() == want}; oracle loop ~{cpu_ms:.0f} ms for B", flush=True)
