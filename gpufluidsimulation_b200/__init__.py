"""gpufluidsimulation_b200 -- BiMocq^2 advection hot path for NVIDIA B200 (sm_100a).

The product is the shared library ``lib/libbimocq_b200.so`` (hand-written CUDA kernels behind the
C ABI declared in ``include/bimocq_b200.h``).  This package is the thin host-side mirror of the
reference's interface for that path (``MapperBaseGPU`` / ``BimocqSolver::advance`` in
``src/bimocq3D``), used by the tests and the benchmark.  There is no CPU fallback: loading fails
loudly when the library has not been built, and every call fails loudly without a CUDA device.
"""
from .capi import BimocqLibraryError, build_library, library_path, load_library  # noqa: F401

__all__ = ["BimocqLibraryError", "build_library", "library_path", "load_library"]
