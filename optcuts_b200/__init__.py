"""optcuts_b200 — B200-native (sm_100a, fp64) implementation of OptCuts' geometry-and-topology
inner loop behind the reference's Energy / LinSysSolver / Optimizer surface.

The product is ``lib/liboptcuts_b200.so`` (C-ABI: ``include/optcuts_b200.h``, built by
``python -m optcuts_b200.build``).  This package is the Python host-side mirror of the reference
interface for that path (same names, argument meaning and error behaviour), used by the parity
tests and by ``bench.py``; the C++ drop-in subclasses are in ``shim/``.
There is NO CPU fallback: every numerical entry point fails loudly if the CUDA library is missing.
"""
from ._capi import Context, OcbError, lib_path, load_library  # noqa: F401
from .trimesh import TriMesh  # noqa: F401
from .energy import SymDirichletEnergy  # noqa: F401
from .linsys import CudaLinSysSolver  # noqa: F401
from .optimizer import Optimizer  # noqa: F401
from . import scaffold, synth  # noqa: F401

__all__ = ["Context", "OcbError", "TriMesh", "SymDirichletEnergy", "CudaLinSysSolver", "Optimizer",
           "lib_path", "load_library"]
