"""Array plumbing shared by the op modules: accept torch tensors (any device)
or NumPy arrays, run on the CUDA device in float32, hand back the caller's
array kind.  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import warnings

import numpy as np
import torch

from . import _lib


class PrecisionWarning(UserWarning):
    pass


_warned_f64 = False


def _warn_f64():
    global _warned_f64
    if not _warned_f64:
        _warned_f64 = True
        warnings.warn(
            "pymotion_b200 computes in float32 on the GPU; float64 inputs are rounded to float32 and the "
            "results are returned as float64 (the NumPy reference would have computed in float64)",
            PrecisionWarning,
            stacklevel=4,
        )


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("pymotion_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class Marshal:
    """Remembers what the first array argument looked like so results can be
    returned the same way (NumPy in -> NumPy out, CPU tensor in -> CPU tensor out)."""

    def __init__(self, *arrays):
        self.kind = "torch"
        self.device = None
        self.out_dtype = torch.float32
        first = True
        for a in arrays:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and self.device is None:
                    self.device = a.device
                if first:
                    self.kind = "torch" if a.is_cuda else "torch_cpu"
                    self.out_dtype = a.dtype if a.dtype.is_floating_point else torch.float32
                    first = False
            elif first and a is not None:
                arr = np.asarray(a)
                self.kind = "numpy"
                self.out_dtype = torch.float64 if arr.dtype == np.float64 else torch.float32
                first = False
        if self.device is None:
            self.device = default_device()
        if self.out_dtype == torch.float64:
            _warn_f64()
        elif self.out_dtype not in (torch.float32,):
            raise TypeError(f"pymotion_b200 supports float32 (and float64 by rounding) inputs, got {self.out_dtype}")

    @staticmethod
    def dev_check(a: torch.Tensor) -> None:
        if a.requires_grad and torch.is_grad_enabled():
            # the reference's torch twins are differentiable (ops/skeleton_torch.py:60-63 clones for exactly
            # that); these kernels have no backward, and handing back a tensor without a graph would train
            # with a silent zero gradient
            raise RuntimeError(
                "pymotion_b200 kernels are forward-only: an input has requires_grad=True. Call under "
                "torch.no_grad() / pass x.detach(), or use the reference's torch twin where gradients are needed")

    def dev(self, a) -> torch.Tensor:
        """float32, on the compute device; NOT necessarily contiguous."""
        if isinstance(a, torch.Tensor):
            self.dev_check(a)
            return a.to(device=self.device, dtype=torch.float32, non_blocking=True)
        return torch.as_tensor(np.asarray(a), device=self.device).to(torch.float32)

    def new(self, shape) -> torch.Tensor:
        return torch.empty(tuple(shape), device=self.device, dtype=torch.float32)

    def out(self, t: torch.Tensor):
        if self.out_dtype != torch.float32:
            t = t.to(self.out_dtype)
        if self.kind == "numpy":
            return t.cpu().numpy()
        if self.kind == "torch_cpu":
            return t.cpu()
        return t

    def out_host(self, t: torch.Tensor):
        """Result that already lives in host memory (the chunked host pipeline)."""
        if self.out_dtype != torch.float32:
            t = t.to(self.out_dtype)
        return t.numpy() if self.kind == "numpy" else t

    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream


def host_parents(parents) -> np.ndarray:
    """parents as a contiguous host int64 array (NumPy, torch on any device, or a list)."""
    if isinstance(parents, torch.Tensor):
        parents = parents.detach().cpu().numpy()
    arr = np.asarray(parents)
    if arr.ndim != 1:
        raise ValueError(f"parents must be one-dimensional, got shape {arr.shape}")
    if arr.dtype.kind not in "iu":
        if arr.dtype.kind == "f" and np.all(arr == np.floor(arr)):
            arr = arr.astype(np.int64)
        else:
            raise ValueError(f"parents must be integers, got dtype {arr.dtype}")
    return np.ascontiguousarray(arr, dtype=np.int64)


def ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


def call(name: str, device: torch.device, *args) -> None:
    lib = _lib.load()
    with torch.cuda.device(device):
        _lib.check(getattr(lib, name)(*args))
