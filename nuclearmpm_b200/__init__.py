"""nuclearmpm_b200 — B200-native MLS-MPM step behind NuclearMPM's embeddable API.

The product is `libnmpm.so` (hand-written sm_100a CUDA behind the C-ABI of include/nmpm.h) plus
include/nclr.h, the drop-in C++ header.  This package is the Python mirror of the same surface
(reference: class MPMSimulation<dim>, src/nclr.h:63-87) used by the tests and bench.py.
There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from .sim import (MPMBatch, MPMSimulation, MaterialModel, NmpmError, OutOfGridError, cube, lib_path, load_library,  # noqa: F401
                  polar_batch, snow_project_batch, svd_batch)

__all__ = ["MPMBatch", "MPMSimulation", "MaterialModel", "NmpmError", "OutOfGridError", "cube", "lib_path", "load_library",
           "svd_batch", "polar_batch", "snow_project_batch"]
