"""pymotion_b200 -- B200-native (sm_100a) drop-in for pymotion's batched forward
kinematics / root dual-quaternion path.

    import pymotion_b200.ops.skeleton as sk          # was: pymotion.ops.skeleton(_torch)
    import pymotion_b200.rotations.quat as quat      # was: pymotion.rotations.quat(_torch)
    import pymotion_b200.rotations.dual_quat as dq   # was: pymotion.rotations.dual_quat(_torch)

Same function names, argument order and return order as the reference; every
call runs one hand-written CUDA kernel through the C ABI in
include/pymotion_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
