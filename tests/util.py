"""Helpers shared by the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def params_from_flat(flat, D, H, device="cpu", dtype=torch.float32):
    """[b1 | W1 | b2 | W2] -> Haiku-layout dict (ravel_pytree order, sorted keys)."""
    flat = torch.as_tensor(flat, dtype=dtype, device=device)
    o1, o2, o3 = H, H + (D + 1) * H, H + (D + 1) * H + D
    return {"mlp_dynamics/linear": {"b": flat[:o1], "w": flat[o1:o2].reshape(D + 1, H)},
            "mlp_dynamics/linear_1": {"b": flat[o2:o3], "w": flat[o3:].reshape(H + 1, D)}}


def flat_from_params(p):
    return torch.cat([p["mlp_dynamics/linear"]["b"].reshape(-1), p["mlp_dynamics/linear"]["w"].reshape(-1),
                      p["mlp_dynamics/linear_1"]["b"].reshape(-1), p["mlp_dynamics/linear_1"]["w"].reshape(-1)])


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
