"""kspace_neutrinos_b200 -- B200-native (sm_100a) build of the per-PM-step k-space hot path of
sbird/kspace-neutrinos behind the reference's C API.

`capi`  : ctypes binding of libkspace_neutrinos_b200.so (the C-ABI in include/*.h)
`host`  : slab partition, grid buffers, the per-process integrator wrapper, collective bootstrap
"""
from . import capi  # noqa: F401

__all__ = ["capi", "host"]
