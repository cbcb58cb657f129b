"""Module-path alias for ``pymotion.ops.vector_torch (/root/reference/pymotion/ops/vector_torch.py)``.

The reference ships a NumPy module and a torch twin with the same function names; here ONE implementation
serves both (tensors in -> tensors out on the same device, NumPy in -> NumPy out), so the twin is the same
module under the twin's name: ``import pymotion_b200.ops.vector_torch`` is the only edit a caller makes."""
from . import vector as _impl
from .vector import *  # noqa: F401,F403

# everything public of the implementation module, including names a star import would skip
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("_")})
