"""rsvd-delta-t_b200 -- B200-native (sm_100a) implementation of RSVD-Delta-t's time-stepping hot path.

The product is the C-ABI shared library `librsvdt_b200.so` (csrc/, include/rsvdt.h) and the C host driver
(host/).  The Python modules here are thin host-side mirrors used by tests and bench.py:

  capi       ctypes binding of include/rsvdt.h (no torch types cross the boundary)
  petsc_io   PETSc-binary readers / writers (the reference's file formats)
  synthetic  synthetic operators / forcings of the BASELINE.json configurations
  dist       row partition over GPUs: halo planning (host arithmetic) and the IPC bootstrap of the CUDA library

The directory name contains '-' (it is the name the task prescribes), so import it through
`__graft_entry__.load_package()` which registers it as the module `rsvdt_b200`.
"""
from . import petsc_io, synthetic, transient_est  # noqa: F401
from . import capi  # noqa: F401
from . import dist  # noqa: F401

__all__ = ["capi", "dist", "petsc_io", "synthetic", "transient_est"]
