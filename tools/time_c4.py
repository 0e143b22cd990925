"""Timing of the C4 forward path (ffjord_mnist.py evaluation closures on the ConcatConv2D kernels): the log-density solve
(ffjord_dynamics) and the forward_all solve (all_aug_dynamics) at the script's network size.
Usage: python tools/time_c4.py [B ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easy_neural_ode_b200 as eno
from easy_neural_ode_b200 import ffjord_mnist_step as fs

dev = torch.device("cuda", 0)
for B in [int(a) for a in sys.argv[1:]] or [25, 200]:
    m = fs.ImageFfjord((28, 28), 64, reg="r2", rtol=1e-5, atol=1e-5, device=dev)
    m.params = m.spec.flatten_params(m.spec.init_params(seed=0, device=dev, scale=1.0, last_scale=1.0)).clone()
    g = torch.Generator().manual_seed(B)
    images = torch.rand((B, 28, 28, 1), generator=g).to(dev)
    eps = (torch.randint(0, 2, (B, 28, 28, 1), generator=g).float() * 2 - 1).to(dev)
    for name, fn in (("ffjord_dynamics (nfe_fn)", lambda: m.nfe(images, eps)), ("all_aug_dynamics (forward_all)", lambda: m.forward_all(images, eps)["nfe"])):
        nfe = fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        streams = 2 if "ffjord" in name else 3          # row streams through the two 64 -> 64 layers per evaluation
        flop = nfe * streams * 2 * (2.0 * B * 784 * 64 * 576)   # algorithmic: interior pixels only, without the time channel
        print(f"B {B:4d} {name:32s}: nfe {nfe:.0f}, {ms:8.2f} ms per solve, {nfe / ms * 1e3:9.0f} NFE/s, "
              f"{flop / ms / 1e9:6.1f} algorithmic TFLOP/s in the tensor-core layers")
