"""resolv_b200 — B200-native (sm_100a CUDA) implementation of SolarLune/resolv's collision hot path.

Only what the path needs: `csrc/` (CUDA kernels + the C ABI of include/resolv_b200.h, built into
libresolv_b200.so), `_abi` (ctypes binding of that ABI), `worlds` (deterministic synthetic worlds) and
`replay` (host-side OnIntersect replay with the reference's ordering). There is no CPU fallback.
"""
from . import _abi, worlds  # noqa: F401
from ._abi import DeviceSpace, RsvError, space_for_world  # noqa: F401

__all__ = ["DeviceSpace", "RsvError", "space_for_world", "worlds"]
