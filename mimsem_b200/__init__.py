"""mimsem_b200 -- B200-native horizontal mixed-mimetic operator path (the hot path of davelee2804/MiMSEM).

The product is the C-ABI shared library ``libmimsem_gpu.so`` (include/mimsem_gpu.h; sources under
``csrc/``) and the C++ host mirror of the reference's classes under ``host/``.  This Python package is
the thin ctypes harness tests and bench.py drive it with; it holds no numerics of its own and has no
CPU fallback: anything that computes raises if the CUDA library or a GPU is missing.
"""
from .lib import load_library, MimsemError  # noqa: F401
from .mesh import Basis, Mesh, patch_topology, write_input  # noqa: F401
from .engine import Engine  # noqa: F401
