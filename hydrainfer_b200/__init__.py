"""hydrainfer_b200 — B200-native (sm_100a) implementation of hydrainfer's paged-KV attention hot path.

Layout:
  csrc/            CUDA kernels + the C ABI of include/hi_b200.h, built in-tree into lib/libhi_b200.so
  csrc/torch_binding.cpp   the compiled Python boundary (pybind11 + ATen over the C ABI), installed as
  _C/              the reference's five pybind modules (hydrainfer._C.kernel.*, hydrainfer._C.data_transfer.*), same names
  _lib.py          ctypes binding of the same C ABI: the no-torch path and the bench / test hooks (fails loudly when the
                   library is missing; there is no CPU fallback)
  dropin.py        registers _C/ under the reference's import names
  memory/, layer/  mirrors of hydrainfer.memory and hydrainfer.layer.causal_attention for this path
"""
__version__ = "0.1.0"
