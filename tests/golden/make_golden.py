"""Generates tests/golden/*.npz by running the UNMODIFIED reference solver
(/root/reference/lib/ode.py) on the torch-backed JAX stand-in (oracle/jax_shim.py).

Run in the build container only (``python tests/golden/make_golden.py``); the
GPU box has no /root/reference and only reads the committed .npz files.

What the vectors pin: the reference's own control flow and arithmetic for the
dopri5 loop, the eight other tableaux loops, NFE accounting, dense output, and
the adjoint (_odeint_rev via jax.vjp -> torch.func.vjp) including the split
wrapper odeint_aux_one.  The dynamics closures (MLP, Taylor-mode regulariser)
come from oracle/dynamics_oracle.py because the reference builds them with
Haiku + jax.experimental.jet, which cannot be imported here.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.jax_shim import load_reference_ode, ravel_pytree  # noqa: E402
from oracle import dynamics_oracle as do  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TABLEAUX = ["heun", "fehlberg", "bosh", "rk_fehlberg", "cash_karp", "owrenzen", "owrenzen5", "tanyam"]


def flat_params(p):
    return ravel_pytree(p)[0].numpy()


def mlp_case(B, D, H, seed, scale):
    g = torch.Generator().manual_seed(1000 + seed)
    params = do.init_mlp_params(D, H, seed=seed, scale=scale)
    x0 = torch.rand((B, D), generator=g)
    eps = torch.randint(0, 2, (B, D), generator=g).float() * 2 - 1
    gy = torch.randn((B, D), generator=g)
    return params, x0, eps, gy


def fin_case(ref, do, B, D, H, params, x0, eps, gy):
    """odeint_sepaux forward + adjoint with the 'fin' aug_dynamics (mnist.py:294-313, lib/ode.py:678-697,
    1686-1693): two tolerances on ts = [0, 1] and a three-point grid with cotangents at every time."""
    rec = dict(B=B, D=D, H=H, x0=x0.numpy(), params=flat_params(params), eps=eps.numpy(), gy=gy.numpy(),
               g_fro=0.01, g_kin=0.02)
    cl = do.make_mnist_closures("none", reg_type="fin")
    y0 = (x0, torch.zeros(B), torch.zeros(B))
    fwd_w = ref.ravel_first_arg(cl["fwd"], ravel_pytree(y0[0])[1])
    rev_w = ref.ravel_first_arg(cl["aug_dynamics"], ravel_pytree(y0)[1])
    for tag, tol, ts in [("5", 1e-5, torch.tensor([0., 1.])), ("d", 1.4e-8, torch.tensor([0., 1.])),
                         ("m", 1e-6, torch.tensor([0., 0.4, 1.]))]:
        (res, nfe), saved = ref._odeint_sepaux_fwd(fwd_w, rev_w, tol, tol, math.inf, y0, ts, eps, params)
        g = tuple(torch.zeros_like(r) for r in res)
        if tag == "m":
            for i in range(len(ts)):
                g[0][i] = gy * (0.5 + i)
                g[1][i] = 0.01 * (1 + i)
                g[2][i] = 0.02 * (3 - i)
        else:
            g[0][-1] = gy
            g[1][-1] = 0.01
            g[2][-1] = 0.02
        out = ref._odeint_sepaux_rev(fwd_w, rev_w, tol, tol, math.inf, saved, (g, None))
        k = f"fin_{tag}"
        rec[f"{k}_ts"], rec[f"{k}_tol"] = ts.numpy(), tol
        rec[f"{k}_y"], rec[f"{k}_nfe"] = res[0].numpy(), float(nfe)
        rec[f"{k}_x0_bar"] = out[0][0].numpy()
        rec[f"{k}_fro0_bar"], rec[f"{k}_kin0_bar"] = out[0][1].numpy(), out[0][2].numpy()
        rec[f"{k}_ts_bar"], rec[f"{k}_eps_bar"] = out[1].numpy(), out[2].numpy()
        rec[f"{k}_params_bar"] = flat_params(out[3])
    np.savez(os.path.join(OUT, "mlp_adjoint_fin.npz"), **rec)


def vmap_case(ref, do, B, D, H, params, x0):
    """Per-example solves (mnist.py:350-353 vmaps odeint over the batch): the reference solver is run on every
    sample separately, which is what each vmap lane computes; plus one mxstep-limited run."""
    cl = do.make_mnist_closures("none")
    ts = torch.tensor([0., 1.])
    rec = dict(B=B, D=D, H=H, x0=x0.numpy(), params=flat_params(params))
    for tag, tol, mx in [("5", 1e-5, math.inf), ("6", 1e-6, math.inf), ("mx2", 1e-5, 2)]:
        ys, nfes = [], []
        for b in range(B):
            y, nfe = ref.odeint(cl["dynamics_wrap"], x0[b:b + 1], ts, params, rtol=tol, atol=tol, mxstep=mx)
            ys.append(y[:, 0].numpy())
            nfes.append(float(nfe))
        rec[f"vmap_{tag}_y"], rec[f"vmap_{tag}_nfe"], rec[f"vmap_{tag}_tol"] = np.stack(ys, 1), np.array(nfes), tol
    np.savez(os.path.join(OUT, "mlp_vmap.npz"), **rec)


def css_case(ref):
    """ffjord_tabular.py's odeint_sepaux call (reg_type 'our': state (y, logp, r)) with ConcatSquash dynamics on
    small dims: groundwork for BASELINE config C2 (oracle/ffjord_oracle.py)."""
    from oracle import ffjord_oracle as fo
    B, D, H, seed, scale = 5, 6, 12, 3, 1.5
    for tol, name in ((1e-5, "css_sepaux.npz"), (1.4e-8, "css_sepaux_default.npz")):   # 1.4e-8: the script default (ffjord_tabular.py:34-35)
        _css_case(ref, fo, B, D, H, seed, scale, tol, name)


def _css_case(ref, fo, B, D, H, seed, scale, tol, name):
    params = fo.init_css_params(D, H, seed=seed, scale=scale)
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn((B, D), generator=g)
    eps = torch.randn((B, D), generator=g)
    gy = torch.randn((B, D), generator=g)
    cl = fo.make_ffjord_closures("r2", "our")
    ts = torch.tensor([0., 1.])
    y0 = (x0, torch.zeros(B, 1), torch.zeros(B, 1))
    fwd_w = ref.ravel_first_arg(cl["fwd"], ravel_pytree(y0[0])[1])
    rev_w = ref.ravel_first_arg(cl["aug_dynamics"], ravel_pytree(y0)[1])
    (res, nfe), saved = ref._odeint_sepaux_fwd(fwd_w, rev_w, tol, tol, math.inf, y0, ts, eps, params)
    gg = tuple(torch.zeros_like(r) for r in res)
    gg[0][-1], gg[1][-1], gg[2][-1] = gy, 1.0 / B, 0.01 / B
    out = ref._odeint_sepaux_rev(fwd_w, rev_w, tol, tol, math.inf, saved, (gg, None))
    np.savez(os.path.join(OUT, name), B=B, D=D, H=H, seed=seed, scale=scale, tol=tol, x0=x0.numpy(),
             eps=eps.numpy(), gy=gy.numpy(), y=res[0].numpy(), nfe=float(nfe), x0_bar=out[0][0].numpy(),
             ts_bar=out[1].numpy(), eps_bar=out[2].numpy(), params_bar=flat_params(out[3]))


def latent_case(ref):
    """One trajectory of latent_ode.py's generative solve (latent_ode.py:344-350): odeint on (z, r) with the R3
    augmented GenDynamics over 5 output times.  Groundwork for BASELINE config C3 (oracle/latent_oracle.py)."""
    from oracle import latent_oracle as lo
    D, U, layers, seed, scale, tol = 8, 12, 2, 4, 1.5, 1e-6
    params = lo.init_gen_params(D, U, layers, seed=seed, scale=scale)
    z0 = 0.5 * torch.randn((1, D), generator=torch.Generator().manual_seed(6))
    ts = torch.tensor([0., 0.2, 0.45, 0.8, 1.0])
    cl = lo.make_latent_closures("r3")
    (zs, rs), nfe = ref.odeint(cl["aug_dynamics"], (z0, torch.zeros(())), ts, params, rtol=tol, atol=tol)
    np.savez(os.path.join(OUT, "latent_gen.npz"), D=D, U=U, layers=layers, seed=seed, scale=scale, tol=tol,
             z0=z0.numpy(), ts=ts.numpy(), zs=zs.numpy(), rs=rs.numpy(), nfe=float(nfe))


def stiff_case(ref):
    """dopri5 forward and adjoint solves that contain REJECTED steps (weights scaled 8x: the controller overshoots
    repeatedly).  Pins the reject branch of the no-clip flavour (lib/ode.py:843-848: old y / f / interpolant kept, dt
    shrunk) on the forward, the dense output after rejections (three output times), and the same for every block of the
    adjoint state including params_bar - for odeint_aux_one (R3) and odeint_sepaux ('fin')."""
    B, D, H, scale = 6, 16, 8, 8.0
    params, x0, eps, gy = mlp_case(B, D, H, seed=0, scale=scale)
    rec = dict(B=B, D=D, H=H, scale=scale, seed=0, x0=x0.numpy(), params=flat_params(params), eps=eps.numpy(), gy=gy.numpy(),
               g_r=0.01, g_fro=0.01, g_kin=0.02)
    cl = do.make_mnist_closures("r3")
    ts3 = torch.tensor([0., 0.35, 1.0])
    ys, nfe = ref.odeint(cl["dynamics_wrap"], x0, ts3, params, rtol=1e-6, atol=1e-6)
    rec["fwd3_ts"], rec["fwd3_y"], rec["fwd3_nfe"], rec["fwd3_tol"] = ts3.numpy(), ys.numpy(), float(nfe), 1e-6
    ts = torch.tensor([0., 1.])
    y0 = (x0, torch.zeros(B))
    fwd_w = ref.ravel_first_arg(cl["fwd"], ravel_pytree(y0[0])[1])
    rev_w = ref.ravel_first_arg(cl["aug_dynamics"], ravel_pytree(y0)[1])
    for tag, tol in [("5", 1e-5), ("6", 1e-6)]:
        (res, nfe), saved = ref._odeint_aux_one_fwd(fwd_w, rev_w, tol, tol, math.inf, y0, ts, eps, params)
        g = (torch.zeros_like(res[0]), torch.zeros_like(res[1]))
        g[0][-1] = gy
        g[1][-1] = 0.01
        out = ref._odeint_aux_one_rev(fwd_w, rev_w, tol, tol, math.inf, saved, (g, None))
        k = f"r3_{tag}"
        rec[f"{k}_tol"], rec[f"{k}_y1"], rec[f"{k}_nfe"] = tol, res[0][-1].numpy(), float(nfe)
        rec[f"{k}_x0_bar"], rec[f"{k}_r0_bar"] = out[0][0].numpy(), out[0][1].numpy()
        rec[f"{k}_ts_bar"], rec[f"{k}_params_bar"] = out[1].numpy(), flat_params(out[3])
    clf = do.make_mnist_closures("none", reg_type="fin")
    y0f = (x0, torch.zeros(B), torch.zeros(B))
    fwd_w = ref.ravel_first_arg(clf["fwd"], ravel_pytree(y0f[0])[1])
    rev_w = ref.ravel_first_arg(clf["aug_dynamics"], ravel_pytree(y0f)[1])
    tol = 1e-5
    (res, nfe), saved = ref._odeint_sepaux_fwd(fwd_w, rev_w, tol, tol, math.inf, y0f, ts, eps, params)
    g = tuple(torch.zeros_like(r) for r in res)
    g[0][-1], g[1][-1], g[2][-1] = gy, 0.01, 0.02
    out = ref._odeint_sepaux_rev(fwd_w, rev_w, tol, tol, math.inf, saved, (g, None))
    rec["fin_tol"], rec["fin_y1"], rec["fin_nfe"] = tol, res[0][-1].numpy(), float(nfe)
    rec["fin_x0_bar"], rec["fin_ts_bar"], rec["fin_eps_bar"] = out[0][0].numpy(), out[1].numpy(), out[2].numpy()
    rec["fin_params_bar"] = flat_params(out[3])
    np.savez(os.path.join(OUT, "mlp_stiff.npz"), **rec)


def main():
    ref = load_reference_ode()
    torch.set_num_threads(8)
    if "--only-stiff" in sys.argv:
        stiff_case(ref)
        return

    # ---- KAT: y' = a + (y - (a t + b))^5, exact y = a t + b (lib/ode.py:1782-1797)
    a, b = 0.2, 3.0
    f = lambda y, t: a + (y - (a * t + b)) ** 5
    ts = torch.tensor([1., 8.])
    y0 = torch.tensor([a * 1. + b])
    rec = {}
    ys, nfe = ref.odeint(f, y0, ts, rtol=1e-5, atol=1e-5)
    rec["dopri5_y"], rec["dopri5_nfe"] = ys.numpy(), float(nfe)
    for name in TABLEAUX:
        ys, nfe = getattr(ref, f"_{name}_odeint")(f, 1e-5, 1e-5, math.inf, y0, ts)
        rec[f"{name}_y"], rec[f"{name}_nfe"] = ys.numpy(), float(nfe)
    np.savez(os.path.join(OUT, "kat_const.npz"), **rec)

    # ---- forward, MLP dynamics, several tolerances and output grids
    B, D, H = 6, 16, 8
    params, x0, eps, gy = mlp_case(B, D, H, seed=0, scale=2.0)
    cl = do.make_mnist_closures("r3")
    rec = dict(B=B, D=D, H=H, x0=x0.numpy(), params=flat_params(params), scale=2.0, seed=0)
    for tag, tol in [("d", 1.4e-8), ("5", 1e-5), ("3", 1e-3)]:
        ts = torch.tensor([0., 1.])
        ys, nfe = ref.odeint(cl["dynamics_wrap"], x0, ts, params, rtol=tol, atol=tol)
        rec[f"fwd_{tag}_y"], rec[f"fwd_{tag}_nfe"], rec[f"fwd_{tag}_tol"] = ys.numpy(), float(nfe), tol
    ts = torch.tensor([0., 0.1, 0.15, 0.7, 1.0, 2.5])
    ys, nfe = ref.odeint(cl["dynamics_wrap"], x0, ts, params, rtol=1e-6, atol=1e-6)
    rec["fwd_multi_ts"], rec["fwd_multi_y"], rec["fwd_multi_nfe"] = ts.numpy(), ys.numpy(), float(nfe)
    # mxstep exhaustion: silent early exit (lib/ode.py:829-831)
    # (target 0.07 sits just past the second step, so the extrapolation stays well conditioned)
    ts = torch.tensor([0., 0.07])
    ys, nfe = ref.odeint(cl["dynamics_wrap"], x0, ts, params, rtol=1e-7, atol=1e-7, mxstep=2)
    rec["fwd_mx2_y"], rec["fwd_mx2_nfe"] = ys.numpy(), float(nfe)
    ts = torch.tensor([0., 1.])
    fun = lambda y, t: cl["dynamics_wrap"](y, t, params).reshape(-1)
    for name in TABLEAUX:
        for tag, tol in [("3", 1e-3), ("5", 1e-5)]:
            if name == "tanyam" and tag == "5":
                continue  # tens of thousands of iterations; tol 1e-3 already pins it
            ys, nfe = getattr(ref, f"_{name}_odeint")(fun, tol, tol, math.inf, x0.reshape(-1), ts)
            rec[f"{name}_{tag}_y"], rec[f"{name}_{tag}_nfe"] = ys.numpy(), float(nfe)
    np.savez(os.path.join(OUT, "mlp_forward.npz"), **rec)

    # ---- odeint_aux_one forward + adjoint, R2 and R3, two tolerances
    rec = dict(B=B, D=D, H=H, x0=x0.numpy(), params=flat_params(params), eps=eps.numpy(), gy=gy.numpy(),
               g_r=0.01)
    ts = torch.tensor([0., 1.])
    y0 = (x0, torch.zeros(B))
    for reg in ["r2", "r3"]:
        cl = do.make_mnist_closures(reg)
        fwd_w = ref.ravel_first_arg(cl["fwd"], ravel_pytree(y0[0])[1])
        rev_w = ref.ravel_first_arg(cl["aug_dynamics"], ravel_pytree(y0)[1])
        for tag, tol in [("5", 1e-5), ("d", 1.4e-8)]:
            (res, nfe), saved = ref._odeint_aux_one_fwd(fwd_w, rev_w, tol, tol, math.inf, y0, ts, eps, params)
            g = (torch.zeros_like(res[0]), torch.zeros_like(res[1]))
            g[0][-1] = gy
            g[1][-1] = 0.01
            out = ref._odeint_aux_one_rev(fwd_w, rev_w, tol, tol, math.inf, saved, (g, None))
            k = f"{reg}_{tag}"
            rec[f"{k}_y1"], rec[f"{k}_nfe"] = res[0][-1].numpy(), float(nfe)
            rec[f"{k}_x0_bar"], rec[f"{k}_r0_bar"] = out[0][0].numpy(), out[0][1].numpy()
            rec[f"{k}_ts_bar"], rec[f"{k}_eps_bar"] = out[1].numpy(), out[2].numpy()
            rec[f"{k}_params_bar"] = flat_params(out[3])
    # augmented FORWARD (evaluation path, mnist.py:333-334): regulariser value
    cl = do.make_mnist_closures("r3")
    y0_all = (x0, torch.zeros(B), torch.zeros(B), torch.zeros(B))
    (ya, ra, fro, kin), nfe = ref.odeint(cl["all_aug_dynamics"], y0_all, ts, eps, params, rtol=1e-5, atol=1e-5)
    rec["all_y1"], rec["all_r"], rec["all_fro"], rec["all_kin"], rec["all_nfe"] = (
        ya[-1].numpy(), ra[-1].numpy(), fro[-1].numpy(), kin[-1].numpy(), float(nfe))
    np.savez(os.path.join(OUT, "mlp_adjoint.npz"), **rec)

    fin_case(ref, do, B, D, H, params, x0, eps, gy)
    vmap_case(ref, do, B, D, H, params, x0)
    css_case(ref)
    latent_case(ref)
    stiff_case(ref)

    # ---- plain adjoint on the pendulum (lib/ode.py:1720-1722, 1765-1774), multi-output
    def pend(y, t, m, g):
        return torch.stack([y[1], -m * y[1] - g * torch.sin(y[0])])
    y0 = torch.tensor([math.pi - 0.1, 0.0])
    ts = torch.linspace(0., 2., 5)
    args = (torch.tensor(0.25), torch.tensor(9.8))
    (ys, nfe), res = ref._odeint_fwd(pend, 1e-6, 1e-6, math.inf, y0, ts, *args)
    g = torch.ones_like(ys)
    out = ref._odeint_rev(pend, 1e-6, 1e-6, math.inf, res, (g, None))
    np.savez(os.path.join(OUT, "pendulum.npz"), ts=ts.numpy(), ys=ys.numpy(), nfe=float(nfe),
             y0_bar=out[0].numpy(), ts_bar=out[1].numpy(), m_bar=float(out[2]), g_bar=float(out[3]))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
