"""Development diagnostics (GPU box): where do CUDA and oracle differ, and is it fp32 noise?"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import mnist_step
from oracle import dynamics_oracle as do, ode_oracle as oo
from tests.util import load, params_from_flat, flat_from_params, rel_err

cuda = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count())
what = sys.argv[1] if len(sys.argv) > 1 else "all"

if what in ("all", "rkf"):
    g = load("mlp_forward.npz")
    B, D, H = int(g["B"]), int(g["D"]), int(g["H"])
    spec = eno.MLPDynamics(D, H)
    pc = params_from_flat(g["params"], D, H)
    x0 = torch.as_tensor(g["x0"])
    for name, tol in [("rk_fehlberg", 1e-5), ("cash_karp", 1e-5)]:
        tr = []
        fun = lambda y, t: do.mlp_dynamics(pc, y, t).reshape(-1)
        ys_o, nfe_o, st = oo.solve(fun, x0.reshape(-1), torch.tensor([0., 1.]), tol, tol, solver=name, trace=tr)
        ys_g, stg, trg = eno.solve_with_trace(spec, name, x0.to(cuda), torch.tensor([0., 1.]),
                                              {k: {kk: vv.to(cuda) for kk, vv in v.items()} for k, v in pc.items()}, tol, tol)
        print(name, "oracle", st, "cuda", stg)
        for i in range(max(len(tr), trg.shape[0])):
            o = tr[i] if i < len(tr) else None
            c = trg[i].tolist() if i < trg.shape[0] else None
            print(i, "O", o, "C", c)

if what in ("all", "step"):
    Bn, D, H, lam = 100, 784, 100, 6e-5
    gen = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (Bn, 28, 28, 1), generator=gen, dtype=torch.uint8)
    labels = torch.randint(0, 10, (Bn,), generator=gen)
    eps = torch.randint(0, 2, (Bn, D), generator=gen).float() * 2 - 1
    p_ode = do.init_mlp_params(D, H, seed=0)
    p_post = do.init_post_params(D, seed=1)
    # "truth": fp64 oracle at tight tolerance
    p64 = {k: {kk: vv.double() for kk, vv in v.items()} for k, v in p_ode.items()}
    post64 = {k: v.double() for k, v in p_post.items()}
    truth = do.train_step_grads(p64, post64, do.pre_ode(images, torch.float64), labels, eps.double(), lam, reg="r3",
                                rtol=1e-10, atol=1e-10)
    tg = flat_from_params(truth["grads"]["ode"])
    for tol in (1e-5, 1.4e-8):
        want = do.train_step_grads(p_ode, p_post, do.pre_ode(images), labels, eps, lam, reg="r3", rtol=tol, atol=tol)
        model = mnist_step.MnistNode(reg="r3", lam=lam, rtol=tol, atol=tol, device=cuda)
        model.load_params(p_ode, p_post)
        out = model.loss_and_grads(images.to(cuda), labels.to(cuda), eps.to(cuda))
        wg = flat_from_params(want["grads"]["ode"])
        print(f"tol {tol}: nfe {float(out['nfe'])}/{want['nfe']} bwd {out['bwd_stats'][0]['n_iters']}/{want['bwd_stats'][0]['n_iters']}")
        print("  y1   cuda-vs-oracle", rel_err(out["y1"], want["y1"]), " cuda-vs-truth", rel_err(out["y1"], truth["y1"]),
              " oracle-vs-truth", rel_err(want["y1"], truth["y1"]))
        print("  grad cuda-vs-oracle", rel_err(out["grads"]["ode"], wg), " cuda-vs-truth", rel_err(out["grads"]["ode"], tg),
              " oracle-vs-truth", rel_err(wg, tg))
        d = (out["grads"]["ode"].cpu() - wg).abs()
        i = int(d.argmax()); P1 = H + (D + 1) * H
        print("  worst index", i, "(b1|W1 ends at", P1, ") cuda", float(out["grads"]["ode"][i]), "oracle", float(wg[i]), "truth", float(tg[i]))
        print("  post_w cuda-vs-oracle", rel_err(out["grads"]["post_w"], want["grads"]["post"]["w"]))
