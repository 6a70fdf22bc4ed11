// Rotary position embedding fused with the KV append: the step in front of the attention kernels.
//
// Replaces apply_rotary_pos_emb (reference csrc/kernel/position_embedding/rope.cu:33-117, called by
// FusedKernelRotaryEmbeddingHandler, hydrainfer/layer/rotary_embedding.py:102-133) and, when slot ids are given, also the
// set_kv_cache launch that follows it in ROPECausalGroupedQueryPageAttention.forward (hydrainfer/model/model_forward.py:
// 81-83 -> causal_attention.py:402): the reference rotates q and k in place (one launch, 2-byte scalar accesses, one CTA
// per token) and then reads k and v again to scatter them into the paged pools (second launch).  Here one launch reads
// q, k and v once as 16-byte vectors, writes the rotated q in place and the rotated k and v straight into their cache
// slots; k is written back only on request (nothing after the append reads it).
//
// Arithmetic is the reference's, bit for bit.  For a pair (x, y) of one head with cos c and sin s of the token's position:
//     x' = x*c - y*s        y' = x*s + y*c
// where every product and the sum are rounded separately - to the element type when the cos/sin table has the element
// type (c10::Half / c10::BFloat16 operators in rope.cu:14-30 and torch's 16-bit tensor ops in
// TorchRotaryEmbeddingHandler.forward, rotary_embedding.py:44-83, both compute in fp32 and round after each operation),
// to fp32 when the table is fp32 (torch promotes q*cos to fp32 and rounds once at the end, :83).  No FMA contraction.
// Pairing: interleaved = (2i, 2i+1); otherwise (i, i + rotary_dim/2).  Dims at or beyond rotary_dim pass through.
#include "common.cuh"

namespace hi {

struct RopeArgs {
  char* q;
  char* k;
  const char* v;
  int64_t q_row_stride, k_row_stride, v_row_stride;  // elements
  const void* positions;
  const char* cos_sin;  // [max_positions, 2, rotary_dim / 2]
  const int32_t* slot_ids;
  char* key_cache;
  char* value_cache;
  int n_qo_heads, n_kv_heads, head_dim, rotary_dim;
  int positions_int64, interleaved, write_back_k;
};

template <typename T, bool kRoundToT>
__device__ __forceinline__ float rope_round(float v) {
  if constexpr (kRoundToT) return Elem<T>::to_f32(Elem<T>::from_f32(v));
  return v;
}

// (x, y) -> (x*c - y*s, x*s + y*c) with the reference's per-operation rounding; results still to be rounded to T.
template <typename T, bool kRoundToT>
__device__ __forceinline__ void rope_pair(float x, float y, float c, float s, float& xo, float& yo) {
  const float xc = rope_round<T, kRoundToT>(__fmul_rn(x, c));
  const float ys = rope_round<T, kRoundToT>(__fmul_rn(y, s));
  const float xs = rope_round<T, kRoundToT>(__fmul_rn(x, s));
  const float yc = rope_round<T, kRoundToT>(__fmul_rn(y, c));
  xo = __fsub_rn(xc, ys);
  yo = __fadd_rn(xs, yc);
}

// 8 consecutive elements <-> fp32 registers, moved as 16-byte vectors.
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]) {
  if constexpr (sizeof(T) == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 w = *reinterpret_cast<const uint4*>(p);
    unpack2<T>(w.x, f[0], f[1]);
    unpack2<T>(w.y, f[2], f[3]);
    unpack2<T>(w.z, f[4], f[5]);
    unpack2<T>(w.w, f[6], f[7]);
  }
}
template <typename T>
struct Pack8 {
  uint4 lo, hi;  // hi only used by 4-byte elements
};
template <typename T>
__device__ __forceinline__ Pack8<T> pack8(const float (&f)[8]) {
  Pack8<T> r;
  if constexpr (sizeof(T) == 4) {
    r.lo = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    r.hi = make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
  } else {
    r.lo = make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7]));
    r.hi = r.lo;
  }
  return r;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const Pack8<T>& v) {
  *reinterpret_cast<uint4*>(p) = v.lo;
  if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4*>(p + 4) = v.hi;
}

// Vector kernel: rotary_dim % 16 == 0, head_dim % 8 == 0, every row 16-byte aligned.  A task is 8 pairs of one head (two
// 8-element vectors in, two out) or one 8-element vector of a pass-through / value copy.  grid = (token, slices of the
// token's task list); small batches get several CTAs per token so a decode step still fills the machine.
template <typename T, typename C>
__global__ void __launch_bounds__(256) rope_append_vec_kernel(const RopeArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr bool kRoundToT = sizeof(T) == 2 && sizeof(C) == 2;
  const int64_t token = blockIdx.x;
  const int n = a.rotary_dim >> 1;
  const int vec_per_half = n >> 3;                        // tasks per head
  const int pass_vecs = (a.head_dim - a.rotary_dim) >> 3; // pass-through vectors per head
  const int head_vecs = a.head_dim >> 3;
  const bool caching = a.slot_ids != nullptr;
  const int n_q = a.n_qo_heads * vec_per_half;
  const int n_k = a.n_kv_heads * vec_per_half;
  const int n_kpass = caching ? a.n_kv_heads * pass_vecs : 0;
  const int n_v = caching ? a.n_kv_heads * head_vecs : 0;
  const int total = n_q + n_k + n_kpass + n_v;

  const int64_t pos = a.positions_int64 ? static_cast<const int64_t*>(a.positions)[token] : static_cast<const int32_t*>(a.positions)[token];
  const C* cos = reinterpret_cast<const C*>(a.cos_sin) + pos * a.rotary_dim;
  const C* sin = cos + n;
  const int64_t slot = caching ? a.slot_ids[token] : 0;
  const int64_t cache_row = static_cast<int64_t>(a.n_kv_heads) * a.head_dim;
  T* q_row = reinterpret_cast<T*>(a.q) + token * a.q_row_stride;
  T* k_row = reinterpret_cast<T*>(a.k) + token * a.k_row_stride;
  T* kc_row = reinterpret_cast<T*>(a.key_cache) + slot * cache_row;

  for (int task = blockIdx.y * blockDim.x + threadIdx.x; task < total; task += gridDim.y * blockDim.x) {
    if (task < n_q + n_k) {
      const bool is_k = task >= n_q;
      const int u = is_k ? task - n_q : task;
      const int h = u / vec_per_half;
      const int p0 = (u - h * vec_per_half) << 3;          // first of the 8 pairs
      T* head = (is_k ? k_row : q_row) + h * a.head_dim;
      const int ix = a.interleaved ? 2 * p0 : p0;          // first vector
      const int iy = a.interleaved ? 2 * p0 + 8 : p0 + n;  // second vector
      float v0[8], v1[8], c[8], s[8];
      load8<T>(head + ix, v0);
      load8<T>(head + iy, v1);
      load8<C>(cos + p0, c);
      load8<C>(sin + p0, s);
      if (a.interleaved) {
        // v0 = (x0 y0 x1 y1 x2 y2 x3 y3), v1 = pairs 4..7
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          rope_pair<T, kRoundToT>(v0[2 * e], v0[2 * e + 1], c[e], s[e], v0[2 * e], v0[2 * e + 1]);
          rope_pair<T, kRoundToT>(v1[2 * e], v1[2 * e + 1], c[4 + e], s[4 + e], v1[2 * e], v1[2 * e + 1]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) rope_pair<T, kRoundToT>(v0[e], v1[e], c[e], s[e], v0[e], v1[e]);
      }
      const Pack8<T> o0 = pack8<T>(v0), o1 = pack8<T>(v1);
      if (!is_k || a.write_back_k) {
        store8<T>(head + ix, o0);
        store8<T>(head + iy, o1);
      }
      if (is_k && caching) {
        T* dst = kc_row + h * a.head_dim;
        store8<T>(dst + ix, o0);
        store8<T>(dst + iy, o1);
      }
    } else if (task < n_q + n_k + n_kpass) {
      const int u = task - n_q - n_k;
      const int h = u / pass_vecs;
      const int e0 = a.rotary_dim + ((u - h * pass_vecs) << 3);
      const T* src = k_row + h * a.head_dim + e0;
      T* dst = kc_row + h * a.head_dim + e0;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
      if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4*>(dst + 4) = *reinterpret_cast<const uint4*>(src + 4);
    } else {
      const int u = task - n_q - n_k - n_kpass;
      const T* src = reinterpret_cast<const T*>(a.v) + token * a.v_row_stride + (static_cast<int64_t>(u) << 3);
      T* dst = reinterpret_cast<T*>(a.value_cache) + slot * cache_row + (static_cast<int64_t>(u) << 3);
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
      if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4*>(dst + 4) = *reinterpret_cast<const uint4*>(src + 4);
    }
  }
}

// Any-shape kernel (odd rotary_dim / 16, unaligned rows): one pair or one copied element per task.
template <typename T, typename C>
__global__ void __launch_bounds__(256) rope_append_scalar_kernel(const RopeArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr bool kRoundToT = sizeof(T) == 2 && sizeof(C) == 2;
  const int64_t token = blockIdx.x;
  const int n = a.rotary_dim >> 1;
  const int pass = a.head_dim - a.rotary_dim;
  const bool caching = a.slot_ids != nullptr;
  const int n_q = a.n_qo_heads * n;
  const int n_k = a.n_kv_heads * n;
  const int n_kpass = caching ? a.n_kv_heads * pass : 0;
  const int n_v = caching ? a.n_kv_heads * a.head_dim : 0;
  const int total = n_q + n_k + n_kpass + n_v;
  const int64_t pos = a.positions_int64 ? static_cast<const int64_t*>(a.positions)[token] : static_cast<const int32_t*>(a.positions)[token];
  const C* cos = reinterpret_cast<const C*>(a.cos_sin) + pos * a.rotary_dim;
  const C* sin = cos + n;
  const int64_t slot = caching ? a.slot_ids[token] : 0;
  const int64_t cache_row = static_cast<int64_t>(a.n_kv_heads) * a.head_dim;
  T* q_row = reinterpret_cast<T*>(a.q) + token * a.q_row_stride;
  T* k_row = reinterpret_cast<T*>(a.k) + token * a.k_row_stride;
  T* kc_row = reinterpret_cast<T*>(a.key_cache) + slot * cache_row;
  for (int task = blockIdx.y * blockDim.x + threadIdx.x; task < total; task += gridDim.y * blockDim.x) {
    if (task < n_q + n_k) {
      const bool is_k = task >= n_q;
      const int u = is_k ? task - n_q : task;
      const int h = u / n;
      const int p = u - h * n;
      T* head = (is_k ? k_row : q_row) + h * a.head_dim;
      const int ix = a.interleaved ? 2 * p : p;
      const int iy = a.interleaved ? 2 * p + 1 : p + n;
      float xo, yo;
      rope_pair<T, kRoundToT>(Elem<T>::to_f32(head[ix]), Elem<T>::to_f32(head[iy]), Elem<C>::to_f32(cos[p]), Elem<C>::to_f32(sin[p]), xo, yo);
      const T xt = Elem<T>::from_f32(xo), yt = Elem<T>::from_f32(yo);
      if (!is_k || a.write_back_k) {
        head[ix] = xt;
        head[iy] = yt;
      }
      if (is_k && caching) {
        kc_row[h * a.head_dim + ix] = xt;
        kc_row[h * a.head_dim + iy] = yt;
      }
    } else if (task < n_q + n_k + n_kpass) {
      const int u = task - n_q - n_k;
      const int h = u / pass;
      const int e = a.rotary_dim + (u - h * pass);
      kc_row[h * a.head_dim + e] = k_row[h * a.head_dim + e];
    } else {
      const int u = task - n_q - n_k - n_kpass;
      reinterpret_cast<T*>(a.value_cache)[slot * cache_row + u] = reinterpret_cast<const T*>(a.v)[token * a.v_row_stride + u];
    }
  }
}

template <typename T, typename C>
static int launch_rope_t(const RopeArgs& a, int64_t n_tokens, bool vec, cudaStream_t stream) {
  const bool caching = a.slot_ids != nullptr;
  const int n = a.rotary_dim / 2;
  const int unit = vec ? 8 : 1;
  int64_t tasks = static_cast<int64_t>(a.n_qo_heads + a.n_kv_heads) * (n / unit);
  if (caching) tasks += static_cast<int64_t>(a.n_kv_heads) * ((a.head_dim - a.rotary_dim) / unit + a.head_dim / unit);
  int threads = static_cast<int>(tasks < 256 ? (tasks + 31) / 32 * 32 : 256);
  if (threads < 32) threads = 32;
  // several CTAs per token while the batch alone cannot fill 148 SMs x 4 CTAs
  int64_t slices = (tasks + threads - 1) / threads;
  const int64_t want = (4 * 148 + n_tokens - 1) / n_tokens;
  if (slices > want) slices = want;
  if (slices < 1) slices = 1;
  const dim3 grid(static_cast<unsigned>(n_tokens), static_cast<unsigned>(slices));
  if (vec) {
    HI_CUDA(launch_pdl(rope_append_vec_kernel<T, C>, grid, threads, 0, stream, a));
  } else {
    HI_CUDA(launch_pdl(rope_append_scalar_kernel<T, C>, grid, threads, 0, stream, a));
  }
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

}  // namespace hi

extern "C" int hi_rope_append(const HiRopeArgs* p, void* stream_) {
  using namespace hi;
  reset_launch_count();
  HI_CHECK_ARG(p != nullptr, "rope_append: null args");
  const HiRopeArgs& r = *p;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int es = dtype_size(r.dtype);
  const int cs = dtype_size(r.cos_sin_dtype);
  HI_CHECK_SUPPORTED(es != 0, "rope_append: unsupported dtype %d", r.dtype);
  HI_CHECK_SUPPORTED(cs != 0 && (r.cos_sin_dtype == r.dtype || r.cos_sin_dtype == HI_F32),
                     "rope_append: the cos/sin table must be fp32 or have the element type (got %d for elements %d)", r.cos_sin_dtype, r.dtype);
  HI_CHECK_ARG(r.n_tokens >= 0, "rope_append: negative n_tokens");
  if (r.n_tokens == 0) return HI_OK;
  HI_CHECK_ARG(r.n_qo_heads > 0 && r.n_kv_heads > 0 && r.head_dim > 0, "rope_append: bad head geometry %d/%d x %d", r.n_qo_heads, r.n_kv_heads, r.head_dim);
  HI_CHECK_ARG(r.rotary_dim >= 0 && r.rotary_dim <= r.head_dim && r.rotary_dim % 2 == 0, "rope_append: rotary_dim %d must be even and <= head_dim %d", r.rotary_dim, r.head_dim);
  HI_CHECK_ARG(r.q && r.k && r.positions && r.cos_sin, "rope_append: null tensor pointer");
  HI_CHECK_ARG(r.q_row_stride >= static_cast<int64_t>(r.n_qo_heads) * r.head_dim && r.k_row_stride >= static_cast<int64_t>(r.n_kv_heads) * r.head_dim,
               "rope_append: row stride smaller than the row (q/k must be contiguous over heads and head_dim, rope.cu:100-101)");
  const bool caching = r.slot_ids != nullptr;
  if (caching) {
    HI_CHECK_ARG(r.v && r.key_cache && r.value_cache, "rope_append: slot ids given without v / caches");
    HI_CHECK_ARG(r.v_row_stride >= static_cast<int64_t>(r.n_kv_heads) * r.head_dim, "rope_append: v row stride smaller than the row");
  }
  HI_DEVICE_GUARD(r.device);
  RopeArgs a{};
  a.q = static_cast<char*>(r.q);
  a.k = static_cast<char*>(r.k);
  a.v = static_cast<const char*>(r.v);
  a.q_row_stride = r.q_row_stride;
  a.k_row_stride = r.k_row_stride;
  a.v_row_stride = r.v_row_stride;
  a.positions = r.positions;
  a.cos_sin = static_cast<const char*>(r.cos_sin);
  a.slot_ids = r.slot_ids;
  a.key_cache = static_cast<char*>(r.key_cache);
  a.value_cache = static_cast<char*>(r.value_cache);
  a.n_qo_heads = r.n_qo_heads;
  a.n_kv_heads = r.n_kv_heads;
  a.head_dim = r.head_dim;
  a.rotary_dim = r.rotary_dim;
  a.positions_int64 = r.positions_int64;
  a.interleaved = r.interleaved;
  a.write_back_k = r.write_back_k || !caching;
  // vector path: whole 16-byte vectors everywhere
  const int vec_elems = 16 / es;
  bool vec = r.rotary_dim % 16 == 0 && r.head_dim % 8 == 0 && aligned_to(r.q, 16) && aligned_to(r.k, 16) && aligned_to(r.cos_sin, 16) &&
             r.q_row_stride % vec_elems == 0 && r.k_row_stride % vec_elems == 0 && ((r.rotary_dim / 2) * cs) % 16 == 0;
  if (caching) vec = vec && aligned_to(r.v, 16) && aligned_to(r.key_cache, 16) && aligned_to(r.value_cache, 16) && r.v_row_stride % vec_elems == 0;
  if (r.force_scalar) vec = false;
  const bool table_f32 = r.cos_sin_dtype == HI_F32;
  switch (r.dtype) {
    case HI_F32: return launch_rope_t<float, float>(a, r.n_tokens, vec, stream);
    case HI_F16: return table_f32 ? launch_rope_t<__half, float>(a, r.n_tokens, vec, stream) : launch_rope_t<__half, __half>(a, r.n_tokens, vec, stream);
    default: return table_f32 ? launch_rope_t<__nv_bfloat16, float>(a, r.n_tokens, vec, stream) : launch_rope_t<__nv_bfloat16, __nv_bfloat16>(a, r.n_tokens, vec, stream);
  }
}
