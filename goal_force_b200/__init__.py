"""goal_force_b200 -- B200-native (sm_100a) implementation of the Goal Force / Wan2.2 DiT denoising hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute on the hot path goes through
the C ABI of libgoalforce_b200.so (include/goalforce_b200.h) via ctypes -- see goal_force_b200.capi.
"""
__version__ = "0.1.0"
