// Split-KV paged attention on CUDA cores — the memory-bound decode kernel, and the any-shape generic path.
//
// Computes what TorchCausalGroupedQueryPageAttentionHandler.forward computes (reference
// hydrainfer/layer/causal_attention.py:307-374): query row t of sequence b attends to the first
// vis = L_b - q_b + i + 1 cached tokens (mask rule :339-342), softmax in fp32, output rounded to the
// input dtype.  It supersedes the reference's fused backends for q_len == 1 rows: the FA2 paged kernel
// (csrc/kernel/flash_attn/src/flash_fwd_launch_template.h:71-105, one CTA per (seq, q-head), no KV
// split) and flashinfer's BatchDecodeWithPagedKVCacheKernel.
//
// Work decomposition (grid = rows x kv-head-groups x kv-chunks):
//   * a CTA owns one query row, G query heads that share one KV head, and one contiguous chunk of the
//     row's visible keys; a KV byte is fetched once per CTA for all G heads;
//   * the CTA is 4 warps = 8 half-warps; a half-warp walks 16-token tiles (one page at block_size 16);
//     its 16 lanes each own D/16 contiguous head dims, so one token row of K or V is one fully
//     coalesced 128-bit-per-lane load (d=128, 16-bit dtype: 256 B = two whole 128-B lines);
//   * Q.K partials of the 16 tokens of a tile are reduced across the 16 lanes with a transposing
//     butterfly (15 shuffles for 16 dot products) that leaves token k's score in lane k, which is
//     exactly the layout the online softmax (warp-shuffle max) and the P.V broadcast need;
//   * half-warp states (m, l, o) are merged through shared memory; if the row was split over chunks the
//     CTA writes an fp32 partial (o, m, l) and merge_partials_kernel does the LSE-weighted reduction.
//
// HBM traffic per row and KV head: 2 * vis * D * sizeof(T) bytes of K and V, read exactly once.
#include <cfloat>
#include <cstdlib>

#include "common.cuh"
#include "attn_common.cuh"

namespace hi {


constexpr int kSimtThreads = 128;
constexpr int kHalfWarps = kSimtThreads / 16;
constexpr unsigned kFull = 0xffffffffu;

// Sequence owning query row t: largest b with q_cu[b] <= t.
__device__ __forceinline__ int find_seq(const int32_t* __restrict__ q_cu, int n_seqs, int t) {
  int lo = 0, hi = n_seqs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(q_cu + mid) <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// E contiguous elements of one lane, as raw 32-bit words.
template <typename T, int E>
struct LaneRow {
  static constexpr int W = E * static_cast<int>(sizeof(T)) / 4;
  static_assert(W >= 2, "a lane owns at least 8 bytes of a row");
  uint32_t w[W];

  __device__ __forceinline__ void load(const T* p) {
    if constexpr (W == 2) {
      const uint2 v = ldg_stream_8(p);
      w[0] = v.x;
      w[1] = v.y;
    } else {
#pragma unroll
      for (int i = 0; i < W / 4; ++i) {
        const uint4 v = ldg_stream_16(reinterpret_cast<const char*>(p) + 16 * i);
        w[4 * i + 0] = v.x;
        w[4 * i + 1] = v.y;
        w[4 * i + 2] = v.z;
        w[4 * i + 3] = v.w;
      }
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < W; ++i) w[i] = 0u;
  }
  __device__ __forceinline__ void to_f32(float (&f)[E]) const {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < E; ++i) f[i] = __uint_as_float(w[i]);
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) unpack2<T>(w[i], f[2 * i], f[2 * i + 1]);
    }
  }
};

// Sum N per-lane values across the 16 lanes of a half-warp so that lane l ends with the total of value l.
template <int N>
__device__ __forceinline__ void transpose_reduce16(float (&v)[16], int lane) {
  if constexpr (N >= 2) {
    constexpr int H = N / 2;
    const bool upper = (lane & H) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float send = upper ? v[i] : v[i + H];
      const float keep = upper ? v[i + H] : v[i];
      v[i] = keep + __shfl_xor_sync(kFull, send, H);
    }
    transpose_reduce16<H>(v, lane);
  }
}

// Merge the 8 half-warp (m, l, o) states of a CTA through shared memory and write the row (or its split-KV partial).
// `scratch` holds kHalfWarps * G * (D + 2) floats.
template <typename T, int D, int G>
__device__ __forceinline__ void merge_half_warps_and_store(const SimtArgs& a, float* scratch, const float (&m)[G],
                                                           const float (&lsum)[G], const float (&o)[G][D / 16], int t, int qh0,
                                                           int chunk) {
  constexpr int E = D / 16;
  const int l16 = threadIdx.x & 15;
  const int hw = threadIdx.x >> 4;
  float* sm_o = scratch;                          // [kHalfWarps][G][D]
  float* sm_m = scratch + kHalfWarps * G * D;     // [kHalfWarps][G]
  float* sm_l = sm_m + kHalfWarps * G;            // [kHalfWarps][G]
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float l = lsum[g];
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) l += __shfl_xor_sync(kFull, l, off);
    if (l16 == 0) {
      sm_m[hw * G + g] = m[g];
      sm_l[hw * G + g] = l;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) sm_o[(hw * G + g) * D + l16 * E + e] = o[g][e];
  }
  __syncthreads();

  for (int idx = threadIdx.x; idx < G * D; idx += kSimtThreads) {
    const int g = idx / D;
    const int d = idx - g * D;
    float mm = -INFINITY;
#pragma unroll
    for (int h = 0; h < kHalfWarps; ++h) mm = fmaxf(mm, sm_m[h * G + g]);
    float osum = 0.f, l = 0.f;
#pragma unroll
    for (int h = 0; h < kHalfWarps; ++h) {
      const float w = fast_exp2(sm_m[h * G + g] - mm);  // mm is finite: the chunk has at least one visible key
      osum = fmaf(w, sm_o[(h * G + g) * D + d], osum);
      l = fmaf(w, sm_l[h * G + g], l);
    }
    const int head = qh0 + g;
    if (a.n_chunks == 1) {
      T* orow = static_cast<T*>(a.out) + static_cast<int64_t>(t) * a.out_row_stride + head * D;
      orow[d] = Elem<T>::from_f32(osum / l);
    } else {
      const int64_t pidx = (static_cast<int64_t>(t) * a.n_qo_heads + head) * a.n_chunks + chunk;
      a.part_o[pidx * D + d] = osum;
      if (d == 0) {
        a.part_ml[pidx * 2 + 0] = mm;
        a.part_ml[pidx * 2 + 1] = l;
      }
    }
  }
}

// out[t, first_elem .. first_elem + n_elems) = 0 by the whole CTA.
template <typename T>
__device__ __forceinline__ void zero_row_heads(const SimtArgs& a, int t, int first_elem, int n_elems) {
  T* orow = static_cast<T*>(a.out) + static_cast<int64_t>(t) * a.out_row_stride + first_elem;
  for (int i = threadIdx.x; i < n_elems; i += blockDim.x) orow[i] = Elem<T>::from_f32(0.f);
}

// OPT: the instance behind mha_varlen_fwd's softcap / sliding-window / alibi arguments (the reference's FlashAttention-2 build has all
// three enabled, flash_api.cpp:93-111, 197-213): scores become softcap * tanh(q.k * scale / softcap), minus slope_h * |i_abs - j|,
// restricted to keys i_abs - window_left <= j <= i_abs + window_right, where i_abs = i + kv_len - q_len is the query's position in key
// coordinates (src/mask.h:54-62, 173-195).  Unsplit (one chunk), one head per CTA.
template <typename T, int D, int G, bool OPT = false>
__global__ void __launch_bounds__(kSimtThreads) paged_attn_simt_kernel(const SimtArgs a) {
  constexpr int E = D / 16;                                        // head dims per lane
  constexpr int W = LaneRow<T, E>::W;                              // 32-bit words per lane per row
  constexpr int TB = (64 / W) < 16 ? (64 / W) : 16;                // tokens loaded per batch (<= 64 raw regs)
  static_assert(16 % TB == 0, "token batch must divide the tile");

  const int t = blockIdx.x;
  const int groups_per_kv = a.group / G;
  const int kvh = blockIdx.y / groups_per_kv;
  const int qh0 = kvh * a.group + (blockIdx.y % groups_per_kv) * G;
  const int chunk = blockIdx.z;

  const int b = find_seq(a.q_cu, a.n_seqs, t);
  const int q_start = __ldg(a.q_cu + b);
  const int q_len = __ldg(a.q_cu + b + 1) - q_start;
  const int kv_len = __ldg(a.kv_cu + b + 1) - __ldg(a.kv_cu + b);
  const int i_abs = kv_len - q_len + (t - q_start);  // the row's position in key coordinates
  int vis = i_abs + 1;                               // keys 0 .. vis-1 are visible to this row
  int first_key = 0;
  if constexpr (OPT) {
    vis = a.window_right < 0 ? kv_len : min(kv_len, i_abs + 1 + a.window_right);
    first_key = a.window_left < 0 ? 0 : max(0, i_abs - a.window_left);
  }
  const int tiles_total = (vis + 15) >> 4;
  const int tile_begin = OPT ? (first_key >> 4) : chunk * a.chunk_tiles;
  if (vis <= first_key) {  // no visible key (kv_len < q_len: malformed metadata; or an empty window): a defined result instead of whatever `out` held
    if (chunk == 0) zero_row_heads<T>(a, t, qh0 * D, G * D);
    return;
  }
  if (tile_begin >= tiles_total) return;  // merge_partials_kernel recomputes the valid chunk count
  const int tile_end = min(tiles_total, tile_begin + a.chunk_tiles);
  const int n_iters = (tile_end - tile_begin + kHalfWarps - 1) / kHalfWarps;
  const int32_t* __restrict__ bt = a.block_tables + __ldg(a.cu_blocks + b);
  pdl_wait();               // q and the appended K / V rows come from kernels in front of this one (the metadata above from a copy)
  pdl_launch_dependents();  // the split merge / the next layer's append may start launching

  const int lane = threadIdx.x & 31;
  const int l16 = lane & 15;
  const int hw = threadIdx.x >> 4;

  const T* __restrict__ kbase = static_cast<const T*>(a.kc) + kvh * D + l16 * E;
  const T* __restrict__ vbase = static_cast<const T*>(a.vc) + kvh * D + l16 * E;

  // This lane's slice of the G query heads, pre-multiplied by scale*log2(e) so scores live in the exp2 domain.
  float qf[G][E];
  {
    const T* qrow = static_cast<const T*>(a.q) + static_cast<int64_t>(t) * a.q_row_stride + qh0 * D + l16 * E;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      LaneRow<T, E> r;
      r.load(qrow + g * D);
      r.to_f32(qf[g]);
#pragma unroll
      for (int e = 0; e < E; ++e) qf[g][e] *= a.scale_log2;
    }
  }

  float m[G], lsum[G], o[G][E];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY;
    lsum[g] = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) o[g][e] = 0.f;
  }

  for (int it = 0; it < n_iters; ++it) {
    const int tile = tile_begin + it * kHalfWarps + hw;
    const int pos = tile * 16 + l16;
    const bool valid = (tile < tile_end) && (pos < vis) && (!OPT || pos >= first_key);
    // Physical slot of this lane's token (block_table[pos / bs] * bs + pos % bs, token_cache_manger.py:126-133).
    int slot = -1;
    if (valid) {
      const int pg = pos / a.block_size;
      slot = __ldg(bt + pg) * a.block_size + (pos - pg * a.block_size);
    }

    // ---- S = q . K^T for the 16 tokens of the tile -------------------------------------------------------
    float acc[G][16];
#pragma unroll
    for (int kb = 0; kb < 16; kb += TB) {
      LaneRow<T, E> raw[TB];
#pragma unroll
      for (int k = 0; k < TB; ++k) {
        const int sk = __shfl_sync(kFull, slot, kb + k, 16);
        raw[k].load(kbase + static_cast<int64_t>(sk < 0 ? 0 : sk) * a.tok_stride);  // masked below if sk < 0
      }
#pragma unroll
      for (int k = 0; k < TB; ++k) {
        float kf[E];
        raw[k].to_f32(kf);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float s = 0.f;
#pragma unroll
          for (int e = 0; e < E; ++e) s = fmaf(qf[g][e], kf[e], s);
          acc[g][kb + k] = s;
        }
      }
    }

    // ---- online softmax; lane k now owns token k ------------------------------------------------------------
    float p[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      transpose_reduce16<16>(acc[g], l16);
      float sc = acc[g][0];
      if constexpr (OPT) {  // scores are in the exp2 domain (q was pre-multiplied by scale * log2 e)
        if (a.softcap_log2 > 0.f) {
          float th;
          asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(sc * a.inv_softcap_log2));
          sc = a.softcap_log2 * th;
        }
        if (a.alibi_slopes != nullptr)
          sc -= __ldg(a.alibi_slopes + b * a.alibi_batch_stride + qh0 + g) * 1.4426950408889634f * fabsf(static_cast<float>(i_abs - pos));
      }
      const float s = valid ? sc : -INFINITY;
      float mx = s;
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, off));
      const float m_new = fmaxf(m[g], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;  // half-warp with no valid token yet
      const float alpha = fast_exp2(m[g] - m_safe);
      p[g] = fast_exp2(s - m_safe);
      lsum[g] = lsum[g] * alpha + p[g];  // lane-partial; summed over lanes once at the end
      m[g] = m_new;
#pragma unroll
      for (int e = 0; e < E; ++e) o[g][e] *= alpha;
    }

    // ---- O += P . V -------------------------------------------------------------------------------------------
#pragma unroll
    for (int kb = 0; kb < 16; kb += TB) {
      LaneRow<T, E> raw[TB];
#pragma unroll
      for (int k = 0; k < TB; ++k) {
        const int sk = __shfl_sync(kFull, slot, kb + k, 16);
        raw[k].load(vbase + static_cast<int64_t>(sk < 0 ? 0 : sk) * a.tok_stride);
        if (sk < 0) raw[k].zero();  // p is 0 there, but 0 * garbage could be NaN
      }
#pragma unroll
      for (int k = 0; k < TB; ++k) {
        float vf[E];
        raw[k].to_f32(vf);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float pk = __shfl_sync(kFull, p[g], kb + k, 16);
#pragma unroll
          for (int e = 0; e < E; ++e) o[g][e] = fmaf(pk, vf[e], o[g][e]);
        }
      }
    }
  }

  __shared__ float scratch[kHalfWarps * G * (D + 2)];
  merge_half_warps_and_store<T, D, G>(a, scratch, m, lsum, o, t, qh0, chunk);
}


// ---- streaming variant: KV pages staged through shared memory with cp.async -------------------------------------------
//
// Same algorithm and thread mapping as paged_attn_simt_kernel, but a lane never holds raw KV in registers: for every
// token of a tile it issues a 16-byte (8-byte for D=64) cp.async from the page straight into a private shared-memory
// slot and later reads that slot back with one LDS.  Each lane only ever reads the bytes it copied itself, so
// cp.async.wait_group is the only synchronisation needed (no barrier, no producer warp).  A half-warp owns one K buffer
// and one V buffer of one tile each; K of tile i+1 is in flight while P.V of tile i is computed, V of tile i while
// Q.K^T of tile i is computed, so every half-warp keeps 4 KiB (d=128) of HBM requests outstanding at all times:
// 64 KiB of shared memory per CTA, 3 CTAs per SM, ~96 KiB in flight per SM against the ~45 KiB Little's law asks for.
template <int BYTES>
__device__ __forceinline__ void cp_async_lane(uint32_t smem_dst, const void* gmem_src, bool pred) {
  const int src_bytes = pred ? BYTES : 0;  // src-size 0: nothing is read, the destination is zero-filled
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_bytes) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_bytes) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T, int E>
__device__ __forceinline__ void lds_lane_row(const uint8_t* p, float (&f)[E]) {
  constexpr int W = E * static_cast<int>(sizeof(T)) / 4;
  uint32_t w[W];
  if constexpr (W == 2) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    w[0] = v.x;
    w[1] = v.y;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    w[0] = v.x;
    w[1] = v.y;
    w[2] = v.z;
    w[3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < W; ++i) unpack2<T>(w[i], f[2 * i], f[2 * i + 1]);
}

template <typename T, int D>
constexpr int stream_smem_bytes() {
  return kHalfWarps * 2 * 16 * 16 * (D / 16) * static_cast<int>(sizeof(T));
}

template <typename T, int D, int G>
__global__ void __launch_bounds__(kSimtThreads) paged_attn_stream_kernel(const SimtArgs a) {
  constexpr int E = D / 16;                               // head dims per lane
  constexpr int EB = E * static_cast<int>(sizeof(T));     // bytes per lane per token row: 16 (d=128) or 8 (d=64)
  static_assert(sizeof(T) == 2 && (EB == 16 || EB == 8), "streaming kernel covers 16-bit dtypes with head_dim 64 / 128");
  constexpr int kRowBytes = 16 * EB;                      // one token row of one KV head
  constexpr int kTileBytes = 16 * kRowBytes;              // one 16-token tile
  static_assert(stream_smem_bytes<T, D>() >= static_cast<int>(kHalfWarps * G * (D + 2) * sizeof(float)),
                "the merge scratch aliases the staging buffers");

  extern __shared__ __align__(16) uint8_t stage[];

  const int t = blockIdx.x;
  const int groups_per_kv = a.group / G;
  const int kvh = blockIdx.y / groups_per_kv;
  const int qh0 = kvh * a.group + (blockIdx.y % groups_per_kv) * G;
  const int chunk = blockIdx.z;

  const int b = find_seq(a.q_cu, a.n_seqs, t);
  const int q_start = __ldg(a.q_cu + b);
  const int q_len = __ldg(a.q_cu + b + 1) - q_start;
  const int kv_len = __ldg(a.kv_cu + b + 1) - __ldg(a.kv_cu + b);
  const int vis = kv_len - q_len + (t - q_start) + 1;
  const int tiles_total = (vis + 15) >> 4;
  const int tile_begin = chunk * a.chunk_tiles;
  if (vis <= 0) {
    if (chunk == 0) zero_row_heads<T>(a, t, qh0 * D, G * D);
    return;
  }
  if (tile_begin >= tiles_total) return;
  const int tile_end = min(tiles_total, tile_begin + a.chunk_tiles);
  const int n_iters = (tile_end - tile_begin + kHalfWarps - 1) / kHalfWarps;
  const int32_t* __restrict__ bt = a.block_tables + __ldg(a.cu_blocks + b);
  pdl_wait();               // q and the appended K / V rows come from kernels in front of this one (the metadata above from a copy)
  pdl_launch_dependents();  // the split merge / the next layer's append may start launching

  const int l16 = threadIdx.x & 15;
  const int hw = threadIdx.x >> 4;
  const uint8_t* kbase = reinterpret_cast<const uint8_t*>(static_cast<const T*>(a.kc) + kvh * D + l16 * E);
  const uint8_t* vbase = reinterpret_cast<const uint8_t*>(static_cast<const T*>(a.vc) + kvh * D + l16 * E);
  const int64_t tok_stride_bytes = a.tok_stride * static_cast<int64_t>(sizeof(T));

  uint8_t* k_buf = stage + hw * 2 * kTileBytes + l16 * EB;  // this lane's column of the half-warp's K tile
  uint8_t* v_buf = k_buf + kTileBytes;
  const uint32_t k_buf_s = static_cast<uint32_t>(__cvta_generic_to_shared(k_buf));
  const uint32_t v_buf_s = k_buf_s + kTileBytes;

  // physical slot of this lane's token in iteration `it` (-1: masked / beyond the chunk)
  auto slot_of = [&](int it) -> int {
    const int tile = tile_begin + it * kHalfWarps + hw;
    const int pos = tile * 16 + l16;
    if (it >= n_iters || tile >= tile_end || pos >= vis) return -1;
    const int pg = pos / a.block_size;
    return __ldg(bt + pg) * a.block_size + (pos - pg * a.block_size);
  };
  auto issue_tile = [&](const uint8_t* base, uint32_t buf_s, int slot_lane) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int sk = __shfl_sync(kFull, slot_lane, k, 16);
      cp_async_lane<EB>(buf_s + k * kRowBytes, base + static_cast<int64_t>(sk < 0 ? 0 : sk) * tok_stride_bytes, sk >= 0);
    }
    cp_async_commit();
  };

  int slot = slot_of(0);
  issue_tile(kbase, k_buf_s, slot);
  issue_tile(vbase, v_buf_s, slot);

  float qf[G][E];
  {
    const T* qrow = static_cast<const T*>(a.q) + static_cast<int64_t>(t) * a.q_row_stride + qh0 * D + l16 * E;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      LaneRow<T, E> r;
      r.load(qrow + g * D);
      r.to_f32(qf[g]);
#pragma unroll
      for (int e = 0; e < E; ++e) qf[g][e] *= a.scale_log2;
    }
  }
  float m[G], lsum[G], o[G][E];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY;
    lsum[g] = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) o[g][e] = 0.f;
  }

  for (int it = 0; it < n_iters; ++it) {
    const bool valid = slot >= 0;
    const int slot_next = slot_of(it + 1);

    // ---- S = q . K^T ----------------------------------------------------------------------------------------------
    cp_async_wait<1>();  // K(it) has landed; V(it) may still be in flight
    float acc[G][16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float kf[E];
      lds_lane_row<T, E>(k_buf + k * kRowBytes, kf);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) s = fmaf(qf[g][e], kf[e], s);
        acc[g][k] = s;
      }
    }
    issue_tile(kbase, k_buf_s, slot_next);  // K(it+1) streams in behind the softmax and P.V of this tile

    // ---- online softmax; lane k owns token k ------------------------------------------------------------------------
    float p[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      transpose_reduce16<16>(acc[g], l16);
      const float s = valid ? acc[g][0] : -INFINITY;
      float mx = s;
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, off));
      const float m_new = fmaxf(m[g], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m[g] - m_safe);
      p[g] = fast_exp2(s - m_safe);
      lsum[g] = lsum[g] * alpha + p[g];
      m[g] = m_new;
#pragma unroll
      for (int e = 0; e < E; ++e) o[g][e] *= alpha;
    }

    // ---- O += P . V ----------------------------------------------------------------------------------------------------
    cp_async_wait<1>();  // V(it) has landed; K(it+1) may still be in flight
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float vf[E];
      lds_lane_row<T, E>(v_buf + k * kRowBytes, vf);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float pk = __shfl_sync(kFull, p[g], k, 16);
#pragma unroll
        for (int e = 0; e < E; ++e) o[g][e] = fmaf(pk, vf[e], o[g][e]);
      }
    }
    issue_tile(vbase, v_buf_s, slot_next);  // V(it+1) streams in behind Q.K^T of the next tile
    slot = slot_next;
  }

  cp_async_wait<0>();
  __syncthreads();  // every lane is done with the staging buffers; reuse them as the merge scratch
  merge_half_warps_and_store<T, D, G>(a, reinterpret_cast<float*>(stage), m, lsum, o, t, qh0, chunk);
}

// LSE-weighted reduction of the per-chunk partials of one (row, head): the split-merge step.  D/4 threads own one
// (row, head); a CTA belongs to one row, finds the row's sequence once and walks the row's heads kMergeThreads / (D/4) at a
// time with a stride of gridDim.y, so a prefill batch (thousands of rows x heads) is one wave of CTAs - one per row - and
// rows the producer wrote directly cost one metadata lookup, while a small decode batch still gets a CTA per 4 heads.
constexpr int kMergeThreads = 128;
template <typename T, int D>
__global__ void __launch_bounds__(kMergeThreads) merge_partials_kernel(const SimtArgs a) {
  constexpr int kLanes = D / 4;
  constexpr int kHeadsPerCta = kMergeThreads / kLanes;
  const int t = blockIdx.x;
  const int b = find_seq(a.q_cu, a.n_seqs, t);
  const int q_start = __ldg(a.q_cu + b);
  const int q_len = __ldg(a.q_cu + b + 1) - q_start;
  const int kv_len = __ldg(a.kv_cu + b + 1) - __ldg(a.kv_cu + b);
  pdl_wait();  // the partials come from the attention kernel in front of this one
  pdl_launch_dependents();
  if (a.direct_tile_tokens > 0) {
    // the producer wrote this row's tile directly when the tile's last row fits in one chunk
    const int i_last = min(q_len, ((t - q_start) / a.direct_tile_tokens + 1) * a.direct_tile_tokens) - 1;
    const int vis_last = kv_len - q_len + i_last + 1;
    if (((vis_last + 15) >> 4) <= (a.direct_tiles > 0 ? a.direct_tiles : a.chunk_tiles)) return;
  }
  const int vis = kv_len - q_len + (t - q_start) + 1;
  const int tiles_total = (vis + 15) >> 4;
  if (tiles_total <= 0) return;  // the producer wrote zeros for a row without visible keys
  // a max_kv_len smaller than the real lengths must not walk into the partials of the neighbouring rows
  const int n_valid = min((tiles_total + a.chunk_tiles - 1) / a.chunk_tiles, a.n_chunks);
  const int d4 = threadIdx.x % kLanes;

  for (int head = blockIdx.y * kHeadsPerCta + threadIdx.x / kLanes; head < a.n_qo_heads; head += gridDim.y * kHeadsPerCta) {
    const int64_t base = (static_cast<int64_t>(t) * a.n_qo_heads + head) * a.n_chunks;
    float mm = -INFINITY;
    for (int c = 0; c < n_valid; ++c) mm = fmaxf(mm, a.part_ml[(base + c) * 2]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float l = 0.f;
    for (int c = 0; c < n_valid; ++c) {
      const float w = fast_exp2(a.part_ml[(base + c) * 2] - mm);
      l = fmaf(w, a.part_ml[(base + c) * 2 + 1], l);
      const float4 v = *reinterpret_cast<const float4*>(a.part_o + (base + c) * D + d4 * 4);
      acc.x = fmaf(w, v.x, acc.x);
      acc.y = fmaf(w, v.y, acc.y);
      acc.z = fmaf(w, v.z, acc.z);
      acc.w = fmaf(w, v.w, acc.w);
    }
    const float inv = 1.f / l;
    T* orow = static_cast<T*>(a.out) + static_cast<int64_t>(t) * a.out_row_stride + head * D + d4 * 4;
    orow[0] = Elem<T>::from_f32(acc.x * inv);
    orow[1] = Elem<T>::from_f32(acc.y * inv);
    orow[2] = Elem<T>::from_f32(acc.z * inv);
    orow[3] = Elem<T>::from_f32(acc.w * inv);
  }
}
template <typename T, int D>
static dim3 merge_grid(const SimtArgs& a) {
  constexpr int kHeadsPerCta = kMergeThreads / (D / 4);
  const int head_groups = (a.n_qo_heads + kHeadsPerCta - 1) / kHeadsPerCta;
  // ~2 waves of 128-thread CTAs at most: many rows -> one CTA per row walking all its heads
  int y = (2 * 148 * 16 + a.n_tokens - 1) / a.n_tokens;
  if (y > head_groups) y = head_groups;
  if (y < 1) y = 1;
  return dim3(a.n_tokens, y);
}

// ---- host side ----------------------------------------------------------------------------------------------------

constexpr int kTargetCtas = 148 * 32;  // enough CTAs that the last partial wave is a few % of the launch
constexpr int kMinChunkTiles = 32;     // never split finer than 512 tokens (graph-timed sweep on B200, ctx 2048 MHA: batch 1 / 4 /
                                       // 8 / 16 run 10.5 / 31.8 / 50.3 / 88.4 us at 32 against 11.1 / 33.1 / 51.0 / 90.5 at 16)

int64_t simt_workspace_bytes(int head_dim) {
  // Partials exist only when n_chunks > 1, i.e. when rows*heads/G < kTargetCtas; then
  // rows*heads*n_chunks <= 2*G*kTargetCtas with G <= 4.
  return static_cast<int64_t>(8) * kTargetCtas * (head_dim + 2) * 4;
}

// Split-merge launch shared with the tile kernel's split-KV mode (a.n_tokens rows, a.n_chunks partials per row and head).
int launch_merge_partials(const SimtArgs& a, int dtype, int head_dim, cudaStream_t stream) {
  if (head_dim == 128 && dtype == HI_BF16) {
    HI_CUDA(launch_pdl(merge_partials_kernel<__nv_bfloat16, 128>, merge_grid<__nv_bfloat16, 128>(a), dim3(kMergeThreads), 0, stream, a));
  } else if (head_dim == 128 && dtype == HI_F16) {
    HI_CUDA(launch_pdl(merge_partials_kernel<__half, 128>, merge_grid<__half, 128>(a), dim3(kMergeThreads), 0, stream, a));
  } else if (head_dim == 256 && dtype == HI_BF16) {
    HI_CUDA(launch_pdl(merge_partials_kernel<__nv_bfloat16, 256>, merge_grid<__nv_bfloat16, 256>(a), dim3(kMergeThreads), 0, stream, a));
  } else if (head_dim == 256 && dtype == HI_F16) {
    HI_CUDA(launch_pdl(merge_partials_kernel<__half, 256>, merge_grid<__half, 256>(a), dim3(kMergeThreads), 0, stream, a));
  } else {
    set_error("merge_partials: unsupported head_dim %d / dtype %d", head_dim, dtype);
    return HI_ERR_UNSUPPORTED;
  }
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

template <typename T, int D, int G>
static int launch_simt_g(const SimtArgs& a, cudaStream_t stream) {
  const dim3 grid(a.n_tokens, a.n_kv_heads * (a.group / G), a.n_chunks);
  constexpr bool kStream = sizeof(T) == 2 && D <= 128;  // 8- or 16-byte lane rows: staged through cp.async
  if constexpr (kStream) {
    constexpr int smem = stream_smem_bytes<T, D>();
    static PerDeviceFlags configured;
    HI_CUDA(configure_dynamic_smem(configured, paged_attn_stream_kernel<T, D, G>, smem));
    timing_mark_start(stream);
    HI_CUDA(launch_pdl(paged_attn_stream_kernel<T, D, G>, grid, dim3(kSimtThreads), smem, stream, a));
  } else {
    timing_mark_start(stream);
    HI_CUDA(launch_pdl(paged_attn_simt_kernel<T, D, G>, grid, dim3(kSimtThreads), 0, stream, a));
  }
  timing_mark_stop(stream);
  note_launch();
  HI_CUDA(cudaGetLastError());
  if (a.n_chunks > 1) {
    HI_CUDA(launch_pdl(merge_partials_kernel<T, D>, merge_grid<T, D>(a), dim3(kMergeThreads), 0, stream, a));
    note_launch();
    HI_CUDA(cudaGetLastError());
  }
  return HI_OK;
}

static int pick_group(int group, int max_g) {
  for (int g = max_g; g > 1; g >>= 1)
    if (group % g == 0) return g;
  return 1;
}

template <typename T, int D>
static int launch_simt_opt(SimtArgs& a, cudaStream_t stream) {
  // score options: one chunk (a window's first tiles are skipped inside the kernel, so the chunk bookkeeping of the merge does not apply),
  // one head per CTA
  a.n_chunks = 1;
  a.chunk_tiles = 1 << 28;
  const dim3 grid(a.n_tokens, a.n_qo_heads, 1);
  timing_mark_start(stream);
  HI_CUDA(launch_pdl(paged_attn_simt_kernel<T, D, 1, true>, grid, dim3(kSimtThreads), 0, stream, a));
  timing_mark_stop(stream);
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

template <typename T, int D>
static int launch_simt_td(SimtArgs& a, const HiAttnArgs& args, cudaStream_t stream) {
  if (args.options != 0) return launch_simt_opt<T, D>(a, stream);
  // fp32 is the reference-parity path only: one head per CTA keeps it within the register budget.
  const int G = pick_group(a.group, sizeof(T) == 4 ? 1 : 4);
  const int64_t ctas_per_chunk = static_cast<int64_t>(a.n_tokens) * a.n_kv_heads * (a.group / G);
  const int max_tiles = (args.max_kv_len + 15) / 16;
  int want_chunks = static_cast<int>((kTargetCtas + ctas_per_chunk - 1) / ctas_per_chunk);
  // 512 tokens for a single row (batch 1 wants every SM it can get), 768 once there are a few rows: graph-timed on B200 at
  // ctx 2048 MHA, batch 4 / 8 / 16 run 28.5 / 47.6 / 86.9 us at 48 tiles against 31.8 / 50.3 / 88.4 at 32, batch 1 10.5 at 32
  // against 12.6 at 48
  int min_chunk_tiles = ctas_per_chunk < 64 ? kMinChunkTiles : kMinChunkTiles + 16;
  if (const char* env = tuning_env("HI_SIMT_MIN_CHUNK_TILES")) min_chunk_tiles = atoi(env) > 0 ? atoi(env) : min_chunk_tiles;  // tuning override
  const int max_chunks = (max_tiles + min_chunk_tiles - 1) / min_chunk_tiles;
  if (want_chunks > max_chunks) want_chunks = max_chunks;
  if (want_chunks < 1) want_chunks = 1;
  a.chunk_tiles = (max_tiles + want_chunks - 1) / want_chunks;
  a.chunk_tiles = (a.chunk_tiles + kHalfWarps - 1) / kHalfWarps * kHalfWarps;  // keep all 8 half-warps busy
  a.n_chunks = (max_tiles + a.chunk_tiles - 1) / a.chunk_tiles;
  if (a.n_chunks < 1) a.n_chunks = 1;
  if (a.n_chunks > 1) {
    const int64_t entries = static_cast<int64_t>(a.n_tokens) * a.n_qo_heads * a.n_chunks;
    const int64_t need = entries * (D + 2) * 4;
    if (args.workspace == nullptr || args.workspace_bytes < need) {
      set_error("paged_attention: workspace of %lld bytes is smaller than the %lld needed for %d KV chunks",
                (long long)args.workspace_bytes, (long long)need, a.n_chunks);
      return HI_ERR_WORKSPACE;
    }
    a.part_o = static_cast<float*>(args.workspace);
    a.part_ml = a.part_o + entries * D;
  }
  switch (G) {
    case 4:
      if constexpr (sizeof(T) == 2) return launch_simt_g<T, D, 4>(a, stream);
    case 2:
      if constexpr (sizeof(T) == 2) return launch_simt_g<T, D, 2>(a, stream);
    default: return launch_simt_g<T, D, 1>(a, stream);
  }
}

template <typename T>
static int launch_simt_t(SimtArgs& a, const HiAttnArgs& args, cudaStream_t stream) {
  switch (args.head_dim) {
    case 64: return launch_simt_td<T, 64>(a, args, stream);
    case 128: return launch_simt_td<T, 128>(a, args, stream);
    case 256: return launch_simt_td<T, 256>(a, args, stream);
    default:
      set_error("paged_attention: head_dim %d not supported (64, 128, 256)", args.head_dim);
      return HI_ERR_UNSUPPORTED;
  }
}

int launch_attn_simt(const HiAttnArgs& args, cudaStream_t stream) {
  SimtArgs a{};
  a.q = args.q;
  a.out = args.out;
  a.kc = args.key_cache;
  a.vc = args.value_cache;
  a.q_row_stride = args.q_row_stride;
  a.out_row_stride = args.out_row_stride;
  a.tok_stride = static_cast<int64_t>(args.n_kv_heads) * args.head_dim;
  a.q_cu = args.q_cu_seq_lens;
  a.kv_cu = args.kv_cu_seq_lens;
  a.block_tables = args.block_tables;
  a.cu_blocks = args.cu_blocks_lens;
  a.n_seqs = args.n_seqs;
  a.n_tokens = args.n_tokens;
  a.n_qo_heads = args.n_qo_heads;
  a.n_kv_heads = args.n_kv_heads;
  a.group = args.n_qo_heads / args.n_kv_heads;
  a.block_size = args.block_size;
  a.scale_log2 = args.softmax_scale * 1.4426950408889634f;
  a.window_left = -1;
  a.window_right = 0;
  if (args.options & HI_ATTN_OPT_WINDOW) {
    a.window_left = args.window_left;
    a.window_right = args.window_right;
  }
  if ((args.options & HI_ATTN_OPT_SOFTCAP) && args.softcap > 0.f) {
    a.softcap_log2 = args.softcap * 1.4426950408889634f;
    a.inv_softcap_log2 = 1.f / a.softcap_log2;
  }
  if ((args.options & HI_ATTN_OPT_ALIBI) && args.alibi_slopes != nullptr) {
    a.alibi_slopes = args.alibi_slopes;
    a.alibi_batch_stride = args.alibi_batch_stride;
  }
  switch (args.dtype) {
    case HI_F32: return launch_simt_t<float>(a, args, stream);
    case HI_F16: return launch_simt_t<__half>(a, args, stream);
    case HI_BF16: return launch_simt_t<__nv_bfloat16>(a, args, stream);
    default: set_error("paged_attention: unsupported dtype %d", args.dtype); return HI_ERR_UNSUPPORTED;
  }
}

}  // namespace hi
