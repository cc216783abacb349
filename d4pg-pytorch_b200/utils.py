"""Tensor helpers with the reference's names (utils.py:4-10), device-aware.

`to_tensor` keeps the reference signature (numpy -> fp32 tensor; `volatile` accepted and
ignored) but places the result on the learner's CUDA device when one exists; `to_numpy`
accepts CUDA tensors."""
import numpy as np
import torch


def default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def to_numpy(var):
    return var.detach().cpu().numpy()


def to_tensor(x, volatile=False, requires_grad=True, dtype=torch.float32, device=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dtype=dtype if isinstance(dtype, torch.dtype) else torch.float32)
    t = t.to(device or default_device())
    if requires_grad and t.is_floating_point():
        t.requires_grad_(True)
    return t
