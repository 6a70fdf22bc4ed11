// KV append: bit-exact row scatter into the paged pools.
//
// Replaces set_kv_cache_kernel (reference csrc/kernel/kv_cache_kernels/kv_cache_kernels.cu:17-58) and
// set_image_cache_kernel (csrc/kernel/cache_kernels/cache_kernels.cu:17-53).  The reference copies one
// 2-byte element per thread per iteration with one CTA per token and recomputes head/offset indices that
// cancel out (dst is simply slot * row_elems + i).  Here the row is moved as 16-byte vectors, K and V go
// in the same launch (blockIdx.y selects the tensor) and a CTA never touches more than one row, so the
// traffic is exactly 2 * row_bytes per (token, tensor): the HBM floor for this op.
#include "common.cuh"

namespace hi {

struct ScatterArgs {
  const int32_t* slot_ids;
  const char* src[2];
  char* dst[2];
  int64_t src_row_stride_bytes[2];
  int64_t row_bytes;
};

template <typename Vec>
__global__ void __launch_bounds__(256) scatter_rows_kernel(ScatterArgs a) {
  pdl_wait();                // k / v (and the slot ids) may come from the kernel in front of this one
  pdl_launch_dependents();   // the attention kernel behind may start its prologue while the rows are copied
  const int64_t token = blockIdx.x;
  const int which = blockIdx.y;
  const int64_t slot = a.slot_ids[token];
  // ternaries instead of a[which]: a dynamically indexed kernel parameter would be copied to local memory
  const char* src_base = which ? a.src[1] : a.src[0];
  char* dst_base = which ? a.dst[1] : a.dst[0];
  const int64_t src_stride = which ? a.src_row_stride_bytes[1] : a.src_row_stride_bytes[0];
  const Vec* __restrict__ src = reinterpret_cast<const Vec*>(src_base + token * src_stride);
  Vec* __restrict__ dst = reinterpret_cast<Vec*>(dst_base + slot * a.row_bytes);
  const int n_vec = static_cast<int>(a.row_bytes / sizeof(Vec));
  for (int i = threadIdx.x; i < n_vec; i += blockDim.x) dst[i] = src[i];
}

// The inverse: out[token] = cache[slot_ids[token]] (image_token_cache[slot_ids, :], parameters_builder.py:48-55).
template <typename Vec>
__global__ void __launch_bounds__(256) gather_rows_kernel(const int32_t* __restrict__ slot_ids, const char* __restrict__ cache,
                                                          char* __restrict__ out, int64_t out_row_stride_bytes, int64_t row_bytes) {
  const int64_t token = blockIdx.x;
  const int64_t slot = slot_ids[token];
  const Vec* __restrict__ src = reinterpret_cast<const Vec*>(cache + slot * row_bytes);
  Vec* __restrict__ dst = reinterpret_cast<Vec*>(out + token * out_row_stride_bytes);
  const int n_vec = static_cast<int>(row_bytes / sizeof(Vec));
  for (int i = threadIdx.x; i < n_vec; i += blockDim.x) dst[i] = src[i];
}

static int launch_scatter(const ScatterArgs& a, int n_tensors, int64_t n_tokens, int device, cudaStream_t stream) {
  if (n_tokens == 0) return HI_OK;
  HI_DEVICE_GUARD(device);
  // Widest vector every row start and the row length are aligned to.
  uintptr_t bits = static_cast<uintptr_t>(a.row_bytes);
  for (int t = 0; t < n_tensors; ++t) {
    bits |= reinterpret_cast<uintptr_t>(a.src[t]) | reinterpret_cast<uintptr_t>(a.dst[t]) |
            static_cast<uintptr_t>(a.src_row_stride_bytes[t]);
  }
  const dim3 grid(static_cast<unsigned>(n_tokens), static_cast<unsigned>(n_tensors));
  auto threads_for = [&](size_t vec) {
    int64_t n = a.row_bytes / static_cast<int64_t>(vec);
    int t = static_cast<int>(n < 256 ? n : 256);
    t = (t + 31) / 32 * 32;
    return t < 32 ? 32 : t;
  };
  if (bits % 16 == 0) {
    HI_CUDA(launch_pdl(scatter_rows_kernel<uint4>, grid, dim3(threads_for(16)), 0, stream, a));
  } else if (bits % 8 == 0) {
    HI_CUDA(launch_pdl(scatter_rows_kernel<uint2>, grid, dim3(threads_for(8)), 0, stream, a));
  } else if (bits % 4 == 0) {
    HI_CUDA(launch_pdl(scatter_rows_kernel<uint32_t>, grid, dim3(threads_for(4)), 0, stream, a));
  } else {
    HI_CUDA(launch_pdl(scatter_rows_kernel<uint16_t>, grid, dim3(threads_for(2)), 0, stream, a));
  }
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

}  // namespace hi

extern "C" int hi_set_kv_cache(const int32_t* slot_ids, const void* keys, const void* values, void* key_cache,
                               void* value_cache, int64_t n_tokens, int64_t row_elems, int64_t key_row_stride,
                               int64_t value_row_stride, int dtype, int device, void* stream) {
  using namespace hi;
  reset_launch_count();
  const int es = dtype_size(dtype);
  HI_CHECK_SUPPORTED(es != 0, "set_kv_cache: unsupported dtype %d", dtype);
  HI_CHECK_ARG(n_tokens >= 0 && row_elems > 0, "set_kv_cache: bad extents n_tokens=%lld row_elems=%lld",
               (long long)n_tokens, (long long)row_elems);
  HI_CHECK_ARG(n_tokens == 0 || (slot_ids && keys && values && key_cache && value_cache), "set_kv_cache: null pointer");
  HI_CHECK_ARG(key_row_stride >= row_elems && value_row_stride >= row_elems,
               "set_kv_cache: row stride smaller than the row (keys/values must be contiguous over heads and head_dim)");
  ScatterArgs a{};
  a.slot_ids = slot_ids;
  a.src[0] = static_cast<const char*>(keys);
  a.src[1] = static_cast<const char*>(values);
  a.dst[0] = static_cast<char*>(key_cache);
  a.dst[1] = static_cast<char*>(value_cache);
  a.src_row_stride_bytes[0] = key_row_stride * es;
  a.src_row_stride_bytes[1] = value_row_stride * es;
  a.row_bytes = row_elems * es;
  return launch_scatter(a, 2, n_tokens, device, static_cast<cudaStream_t>(stream));
}

extern "C" int hi_set_image_cache(const int32_t* slot_ids, const void* image_tokens, void* image_cache,
                                  int64_t n_tokens, int64_t row_elems, int64_t token_row_stride, int dtype, int device,
                                  void* stream) {
  using namespace hi;
  reset_launch_count();
  const int es = dtype_size(dtype);
  HI_CHECK_SUPPORTED(es != 0, "set_image_cache: unsupported dtype %d", dtype);
  HI_CHECK_ARG(n_tokens >= 0 && row_elems > 0, "set_image_cache: bad extents");
  HI_CHECK_ARG(n_tokens == 0 || (slot_ids && image_tokens && image_cache), "set_image_cache: null pointer");
  HI_CHECK_ARG(token_row_stride >= row_elems, "set_image_cache: row stride smaller than the row");
  ScatterArgs a{};
  a.slot_ids = slot_ids;
  a.src[0] = a.src[1] = static_cast<const char*>(image_tokens);
  a.dst[0] = a.dst[1] = static_cast<char*>(image_cache);
  a.src_row_stride_bytes[0] = a.src_row_stride_bytes[1] = token_row_stride * es;
  a.row_bytes = row_elems * es;
  return launch_scatter(a, 1, n_tokens, device, static_cast<cudaStream_t>(stream));
}

extern "C" int hi_get_image_cache(const int32_t* slot_ids, const void* image_cache, void* out, int64_t n_tokens,
                                  int64_t row_elems, int64_t out_row_stride, int dtype, int device, void* stream_) {
  using namespace hi;
  reset_launch_count();
  const int es = dtype_size(dtype);
  HI_CHECK_SUPPORTED(es != 0, "get_image_cache: unsupported dtype %d", dtype);
  HI_CHECK_ARG(n_tokens >= 0 && row_elems > 0, "get_image_cache: bad extents");
  HI_CHECK_ARG(n_tokens == 0 || (slot_ids && image_cache && out), "get_image_cache: null pointer");
  HI_CHECK_ARG(out_row_stride >= row_elems, "get_image_cache: row stride smaller than the row");
  if (n_tokens == 0) return HI_OK;
  HI_DEVICE_GUARD(device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t row_bytes = row_elems * es, stride_bytes = out_row_stride * es;
  const uintptr_t bits = static_cast<uintptr_t>(row_bytes) | static_cast<uintptr_t>(stride_bytes) |
                         reinterpret_cast<uintptr_t>(image_cache) | reinterpret_cast<uintptr_t>(out);
  const char* src = static_cast<const char*>(image_cache);
  char* dst = static_cast<char*>(out);
  const unsigned grid = static_cast<unsigned>(n_tokens);
  auto threads_for = [&](size_t vec) {
    const int64_t n = row_bytes / static_cast<int64_t>(vec);
    int t = static_cast<int>(n < 256 ? n : 256);
    t = (t + 31) / 32 * 32;
    return t < 32 ? 32 : t;
  };
  if (bits % 16 == 0) {
    gather_rows_kernel<uint4><<<grid, threads_for(16), 0, stream>>>(slot_ids, src, dst, stride_bytes, row_bytes);
  } else if (bits % 8 == 0) {
    gather_rows_kernel<uint2><<<grid, threads_for(8), 0, stream>>>(slot_ids, src, dst, stride_bytes, row_bytes);
  } else if (bits % 4 == 0) {
    gather_rows_kernel<uint32_t><<<grid, threads_for(4), 0, stream>>>(slot_ids, src, dst, stride_bytes, row_bytes);
  } else {
    gather_rows_kernel<uint16_t><<<grid, threads_for(2), 0, stream>>>(slot_ids, src, dst, stride_bytes, row_bytes);
  }
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}
