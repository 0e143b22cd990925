"""Generates tests/golden/mnist_trained_step19.npz: the parameters of the MNIST NODE (BASELINE configs[0]) after 19
momentum updates of the ORACLE training trajectory that bench.py times (steps 0..18 from the seeded initial parameters,
synthetic batches of bench.synth_batch), plus the oracle's results for step 19 on those parameters.

Why: every other parity case starts from initial weights.  After 19 updates the dynamics are stiffer (more forward and
adjoint iterations) - this fixture lets the GPU path be checked on exactly the inputs the CPU arm sees at the end of the
benchmarked trajectory (tests/test_gpu_adjoint.py:test_trained_parameters_step).

Run in the build container: ``python tests/golden/make_trained.py`` (about a minute of CPU).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import dynamics_oracle as do  # noqa: E402
from tests.util import flat_from_params  # noqa: E402

N_UPDATES = 19


def main():
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    p_ode = do.init_mlp_params(bench.D, bench.H, seed=0)
    p_post = do.init_post_params(bench.D, seed=1)
    vel = None
    for s in range(N_UPDATES):
        images, labels, eps = bench.synth_batch(0, s)
        out = do.train_step_grads(p_ode, p_post, do.pre_ode(images), labels, eps, bench.LAM, reg="r3", rtol=bench.TOL, atol=bench.TOL)
        params = {"ode": p_ode, "post": p_post}
        if vel is None:
            vel = do.ode_oracle_tree_zeros(params)
        params, vel = do.momentum_update(params, vel, out["grads"], 1e-1)
        p_ode, p_post = params["ode"], params["post"]
        print(s, float(out["loss"]), out["fwd_stats"]["n_iters"], out["bwd_stats"][0]["n_iters"], flush=True)
    images, labels, eps = bench.synth_batch(0, N_UPDATES)
    out = do.train_step_grads(p_ode, p_post, do.pre_ode(images), labels, eps, bench.LAM, reg="r3", rtol=bench.TOL, atol=bench.TOL)
    print("step", N_UPDATES, float(out["loss"]), out["fwd_stats"], out["bwd_stats"][0])
    np.savez_compressed(
        os.path.join(os.path.dirname(os.path.abspath(__file__)), "mnist_trained_step19.npz"),
        step=N_UPDATES, p_ode=flat_from_params(p_ode).numpy(), post_w=p_post["w"].numpy(), post_b=p_post["b"].numpy(),
        loss=float(out["loss"]), nfe=float(out["nfe"]), fwd_iters=out["fwd_stats"]["n_iters"],
        fwd_accepted=out["fwd_stats"]["n_accepted"], bwd_iters=out["bwd_stats"][0]["n_iters"],
        bwd_accepted=out["bwd_stats"][0]["n_accepted"], y1_head=out["y1"][:4].numpy(),
        grad_post_w=out["grads"]["post"]["w"].numpy(), grad_ode_norm=float(flat_from_params(out["grads"]["ode"]).norm()),
        grad_ode_head=flat_from_params(out["grads"]["ode"])[:4096].numpy())


if __name__ == "__main__":
    main()
