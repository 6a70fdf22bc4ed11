"""hydrainfer_b200 — B200-native (sm_100a) implementation of hydrainfer's paged-KV attention hot path.

Layout:
  csrc/            CUDA kernels + the C ABI of include/hi_b200.h, built in-tree into lib/libhi_b200.so
  _lib.py          ctypes binding (fails loudly when the library is missing; there is no CPU fallback)
  _C/              mirrors of the reference's pybind modules (hydrainfer._C.kernel.*, hydrainfer._C.data_transfer.*)
  memory/, layer/  mirrors of hydrainfer.memory and hydrainfer.layer.causal_attention for this path
"""
__version__ = "0.1.0"
