"""Operator boundary kept for extenders: `to_tensor`, `mat_mul`, `elem_mul`.

Same signatures as /root/reference/pyrhe/src/util/mat_mul.py:4-48 (`*mats, device,
to_numpy=True`; numpy inputs become float32 tensors).  The built-in models do not go
through these per-call round trips -- they use the fused block kernels of libpyrhe_b200
(pyrhe_b200/engine.py); these wrappers exist so custom `Base` subclasses written against
the reference keep working, and they refuse to run anywhere but on a CUDA device.
"""
import numpy as np
import torch


def _cuda_device(device):
    if not torch.cuda.is_available():
        raise RuntimeError("pyrhe_b200 has no CPU path: a CUDA device (B200) is required")
    if device is None or getattr(device, "type", str(device)) == "cpu" or device == "cpu":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def to_tensor(x, device=None):
    device = _cuda_device(device)
    if isinstance(x, torch.Tensor):
        return x.to(device=device)
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).to(device=device, dtype=torch.float32)
    raise ValueError(f"Failed to convert {x} to tensor: {x} is neither a tensor or a Numpy array. ")


def _chain(op, mats, device, to_numpy):
    if not mats:
        raise ValueError("At least one matrix is required.")
    out = to_tensor(mats[0], device)
    for m in mats[1:]:
        out = op(out, to_tensor(m, device))
    return out.cpu().numpy() if to_numpy else out


def mat_mul(*mats, device, to_numpy=True):
    return _chain(torch.matmul, mats, device, to_numpy)


def elem_mul(*mats, device, to_numpy=True):
    return _chain(torch.mul, mats, device, to_numpy)
