// Grouped-query decode attention on tcgen05 with the operands swapped ("tokens on the M side").
//
// For one query row of a grouped model the useful work per KV tile is tiny (G <= 16 heads x 128 tokens), so the prefill
// tile kernel — rows = query heads, 128 TMEM lanes of which G are real — wastes the softmax warps on padding.  Here the
// MMA is transposed so that the 128 TMEM lanes are the 128 TOKENS of the KV tile:
//
//     S^T[token, head]  = K_tile[token, :] . Q^T[:, head]          M = 128 tokens, N = 16 (heads, padded), K = 128 dims
//     O^T[dim,   head] += V_tile^T[dim, token] . P^T[token, head]  M = 128 dims,   N = 16,                 K = 128 tokens
//
// K and V tiles are staged exactly as in attn_tc.cu (one TMA box per page and 64-dim half, 128B-swizzled): K is the
// K-major A operand of the first product, V — untouched — the MN-major A operand of the second.  Every softmax thread
// owns one token: it reads its G scores from TMEM, the tile max per head is a warp shuffle + a 4-warp exchange through
// shared memory, and it writes its G probabilities into the K-major P^T operand in shared memory.  Exponentials, masks
// and row sums are G per thread per tile (instead of 128), so the kernel is bound by the TMA stream of KV pages:
// 64 KiB per tile per CTA, three CTAs per SM, split-KV over the grid for load balance with the same fp32 partial + merge
// as the CUDA-core kernel.  Same math as TorchCausalGroupedQueryPageAttentionHandler (reference
// hydrainfer/layer/causal_attention.py:307-374); it replaces flashinfer's BatchDecodeWithPagedKVCacheWrapper
// (use_tensor_cores=True, reference executor.py:100-101) for grouped models.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "attn_common.cuh"
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tma_maps.h"

namespace hi {

constexpr int kDecThreads = 192;
constexpr int kDecTile = 128;                 // tokens per KV tile == MMA M
constexpr int kDecN = 16;                     // MMA N: query heads of one KV head, padded to the minimum for M = 128
constexpr int kDecD = 128;                    // head dim
constexpr int kDecHalf = kDecTile * 128;      // one 64-dim half of a K or V tile: 16 KiB
constexpr int kDecTileBytes = 2 * kDecHalf;   // 32 KiB
constexpr int kDecQHalf = kDecN * 128;        // one 64-dim half of Q^T: 16 rows x 128 B
constexpr int kDecPHalf = kDecN * 128;        // one 64-token half of P^T: 16 rows x 128 B
constexpr uint32_t kDecTmemCols = 32;
constexpr uint32_t kDecColS = 0;
constexpr uint32_t kDecColO = 16;
constexpr float kDecRescale = 8.0f;

struct DecArgs {
  void* out;
  int64_t out_row_stride;
  const int32_t* q_cu;
  const int32_t* kv_cu;
  const int32_t* block_tables;
  const int32_t* cu_blocks;
  int n_seqs, n_qo_heads, n_kv_heads, group, block_size;
  float scale_log2;
  int n_splits, tiles_per_split;
  float* part_o;
  float* part_ml;
};

template <int NST>
struct DecSmem {
  static constexpr int kK = 0;
  static constexpr int kV = NST * kDecTileBytes;
  static constexpr int kQ = 2 * NST * kDecTileBytes;
  static constexpr int kP = kQ + 2 * kDecQHalf;
  static constexpr int kRed = kP + 2 * kDecPHalf;       // [2 parities][4 warps][16 heads] floats
  static constexpr int kBars = kRed + 2 * 4 * 16 * 4;
  static constexpr int bQFull = 0;
  static constexpr int bKFull = 1;
  static constexpr int bKEmpty = 1 + NST;
  static constexpr int bVFull = 1 + 2 * NST;
  static constexpr int bVEmpty = 1 + 3 * NST;
  static constexpr int bSFull = 1 + 4 * NST;
  static constexpr int bPFull = 2 + 4 * NST;
  static constexpr int bOFull = 3 + 4 * NST;
  static constexpr int kNumBars = 4 + 4 * NST;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemPtr + 16;
  static constexpr int kDynamicBytes = kTotal + 1024;
};

__device__ __forceinline__ int dec_find_seq(const int32_t* __restrict__ q_cu, int n_seqs, int t) {
  int lo = 0, hi = n_seqs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(q_cu + mid) <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// GP = heads handled per thread (group size rounded up to 8 or 16).
template <typename T, int NST, int GP>
__global__ void __launch_bounds__(kDecThreads, NST == 1 ? 3 : 1)
paged_decode_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                       const __grid_constant__ CUtensorMap tm_v, const DecArgs a) {
  using L = DecSmem<NST>;
  constexpr bool kBf16 = !std::is_same<T, __half>::value;

  const int t = blockIdx.x;       // query row
  const int kvh = blockIdx.y;
  const int sp = blockIdx.z;
  const int b = dec_find_seq(a.q_cu, a.n_seqs, t);
  const int q_start = __ldg(a.q_cu + b);
  const int q_len = __ldg(a.q_cu + b + 1) - q_start;
  const int kv_len = __ldg(a.kv_cu + b + 1) - __ldg(a.kv_cu + b);
  const int vis = kv_len - q_len + (t - q_start) + 1;  // keys [0, vis) are visible to this row
  const int n_tiles_all = (vis + kDecTile - 1) / kDecTile;
  const int j_begin = sp * a.tiles_per_split;
  if (j_begin >= n_tiles_all) return;
  const int n_tiles = min(n_tiles_all - j_begin, a.tiles_per_split);
  const int blk0 = __ldg(a.cu_blocks + b);
  const int n_pages = __ldg(a.cu_blocks + b + 1) - blk0;
  const int pages_per_tile = kDecTile / a.block_size;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  auto bar = [&](int idx) -> uint32_t { return smem_base + L::kBars + idx * 8; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kTmemPtr);
  float* red = reinterpret_cast<float*>(smem_gen + L::kRed);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar(L::bQFull), 1);
    for (int s = 0; s < NST; ++s) {
      ptx::mbar_init(bar(L::bKFull + s), 1);
      ptx::mbar_init(bar(L::bKEmpty + s), 1);
      ptx::mbar_init(bar(L::bVFull + s), 1);
      ptx::mbar_init(bar(L::bVEmpty + s), 1);
    }
    ptx::mbar_init(bar(L::bSFull), 1);
    ptx::mbar_init(bar(L::bPFull), kDecTile);
    ptx::mbar_init(bar(L::bOFull), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      ptx::prefetch_tensormap(&tm_q);
      ptx::prefetch_tensormap(&tm_k);
      ptx::prefetch_tensormap(&tm_v);
    }
    __syncwarp();
    ptx::tmem_alloc(smem_base + L::kTmemPtr, kDecTmemCols);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // The prologue above (barriers, TMEM, tensor maps; metadata from a copy) ran beside the tail of the append kernel; q and the
  // appended K / V rows are touched only from here on.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 4) {
    // ================================================ TMA producer ================================================
    if (lane == 0) {
      // Q^T operand: rows = the `group` query heads of this KV head (rows group..15 stay unwritten: they only feed
      // output columns nobody reads), one box per 64-dim half.
      ptx::mbar_arrive_expect_tx(bar(L::bQFull), 2u * static_cast<uint32_t>(a.group) * 128u);
      ptx::tma_load_3d(smem_base + L::kQ, &tm_q, bar(L::bQFull), 0, kvh * a.group, t);
      ptx::tma_load_3d(smem_base + L::kQ + kDecQHalf, &tm_q, bar(L::bQFull), 64, kvh * a.group, t);
    }
    const uint32_t page_half_bytes = static_cast<uint32_t>(a.block_size) * 128u;
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j % NST;
      const uint32_t ph = static_cast<uint32_t>(j / NST) & 1u;
      const int page0 = (j_begin + j) * pages_per_tile;
      const int n_valid = min(pages_per_tile, n_pages - page0);
      int blk = 0;
      if (lane < n_valid) blk = __ldg(a.block_tables + blk0 + page0 + lane);
      const uint32_t tx = static_cast<uint32_t>(n_valid > 0 ? n_valid : 0) * 2u * page_half_bytes;
      if (lane == 0) {
        ptx::mbar_wait(bar(L::bKEmpty + st), ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar(L::bKFull + st), tx);
      }
      __syncwarp();
      if (lane < n_valid) {
        const uint32_t dst = smem_base + L::kK + st * kDecTileBytes + lane * page_half_bytes;
        ptx::tma_load_3d(dst, &tm_k, bar(L::bKFull + st), 0, kvh, blk * a.block_size);
        ptx::tma_load_3d(dst + kDecHalf, &tm_k, bar(L::bKFull + st), 64, kvh, blk * a.block_size);
      }
      if (lane == 0) {
        ptx::mbar_wait(bar(L::bVEmpty + st), ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar(L::bVFull + st), tx);
      }
      __syncwarp();
      if (lane < n_valid) {
        const uint32_t dst = smem_base + L::kV + st * kDecTileBytes + lane * page_half_bytes;
        ptx::tma_load_3d(dst, &tm_v, bar(L::bVFull + st), 0, kvh, blk * a.block_size);
        ptx::tma_load_3d(dst + kDecHalf, &tm_v, bar(L::bVFull + st), 64, kvh, blk * a.block_size);
      }
    }
  } else if (warp == 5) {
    // ================================================ MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_f16(kBf16, false, false, kDecTile, kDecN);  // K (K-major) x Q^T (K-major)
      constexpr uint32_t idesc_o = ptx::make_idesc_f16(kBf16, true, false, kDecD, kDecN);      // V^T (MN-major) x P^T (K-major)
      const uint32_t tmem_s = tmem_base + kDecColS;
      const uint32_t tmem_o = tmem_base + kDecColO;
      const uint32_t q_addr = smem_base + L::kQ;
      const uint32_t p_addr = smem_base + L::kP;
      ptx::mbar_wait(bar(L::bQFull), 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % NST;
        const uint32_t ph = static_cast<uint32_t>(j / NST) & 1u;
        ptx::mbar_wait(bar(L::bKFull + st), ph);
        ptx::tc_fence_after_sync();
        const uint32_t k_addr = smem_base + L::kK + st * kDecTileBytes;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 dims per step
          ptx::mma_f16_ss(tmem_s, ptx::make_smem_desc_sw128(k_addr + (kk >> 2) * kDecHalf + (kk & 3) * 32, 16, 1024),
                          ptx::make_smem_desc_sw128(q_addr + (kk >> 2) * kDecQHalf + (kk & 3) * 32, 16, 1024), idesc_s, kk > 0);
        }
        ptx::mma_commit(bar(L::bKEmpty + st));
        ptx::mma_commit(bar(L::bSFull));
        ptx::mbar_wait(bar(L::bPFull), static_cast<uint32_t>(j) & 1u);
        ptx::mbar_wait(bar(L::bVFull + st), ph);
        ptx::tc_fence_after_sync();
        const uint32_t v_addr = smem_base + L::kV + st * kDecTileBytes;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 tokens per step
          ptx::mma_f16_ss(tmem_o, ptx::make_smem_desc_sw128(v_addr + kk * 2048, kDecHalf, 1024),
                          ptx::make_smem_desc_sw128(p_addr + (kk >> 2) * kDecPHalf + (kk & 3) * 32, 16, 1024), idesc_o,
                          (j > 0) || (kk > 0));
        }
        ptx::mma_commit(bar(L::bVEmpty + st));
        if (j == n_tiles - 1) ptx::mma_commit(bar(L::bOFull));
      }
    }
  } else {
    // ================================================ softmax (thread = token) + epilogue (thread = head dim) =======
    const int r = threadIdx.x;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + kDecColS;
    const uint32_t tmem_o = tmem_base + lane_base + kDecColO;
    uint8_t* p_row = smem_gen + L::kP + (r >> 6) * kDecPHalf + (r & 7) * 2;  // + h*128 + (((r&63)>>3) ^ (h&7))*16
    const int p_chunk = (r & 63) >> 3;
    float m_ref[GP], l_thr[GP];
#pragma unroll
    for (int h = 0; h < GP; ++h) {
      m_ref[h] = 0.f;
      l_thr[h] = 0.f;
    }

    for (int j = 0; j < n_tiles; ++j) {
      const int kv0 = (j_begin + j) * kDecTile;
      const bool valid = kv0 + r < vis;
      float* red_j = red + (j & 1) * 64;
      ptx::mbar_wait(bar(L::bSFull), static_cast<uint32_t>(j) & 1u);
      ptx::tc_fence_after_sync();
      uint32_t sv[16];
      ptx::tmem_ld_x16(tmem_s, sv);
      ptx::tmem_wait_ld();
      float s[GP];
#pragma unroll
      for (int h = 0; h < GP; ++h) {
        s[h] = valid ? __uint_as_float(sv[h]) * a.scale_log2 : -INFINITY;
        float mx = s[h];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (lane == 0) red_j[warp * 16 + h] = mx;
      }
      ptx::named_bar_sync(1, kDecTile);
      bool grow = false;
      float mt[GP];
#pragma unroll
      for (int h = 0; h < GP; ++h) {
        mt[h] = fmaxf(fmaxf(red_j[h], red_j[16 + h]), fmaxf(red_j[32 + h], red_j[48 + h]));
        if (h < a.group) grow = grow || (mt[h] > m_ref[h] + kDecRescale);
      }
      if (j == 0) {
#pragma unroll
        for (int h = 0; h < GP; ++h) m_ref[h] = (mt[h] == -INFINITY) ? 0.f : mt[h];
      } else if (grow) {  // uniform over the CTA: every thread sees the same tile maxima
        uint32_t ov[16];
        ptx::tmem_ld_x16(tmem_o, ov);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int h = 0; h < GP; ++h) {
          const float m_new = fmaxf(m_ref[h], mt[h]);
          const float alpha = fast_exp2(m_ref[h] - m_new);
          l_thr[h] *= alpha;
          m_ref[h] = m_new;
          ov[h] = __float_as_uint(__uint_as_float(ov[h]) * alpha);
        }
        ptx::tmem_st_x16(tmem_o, ov);
      }
#pragma unroll
      for (int h = 0; h < GP; ++h) {
        const float p = fast_exp2(s[h] - m_ref[h]);  // masked tokens: exp2(-inf) = 0
        l_thr[h] += p;
        if (h < a.group) *reinterpret_cast<T*>(p_row + h * 128 + ((p_chunk ^ (h & 7)) << 4)) = Elem<T>::from_f32(p);
      }
      // Tokens at or beyond kv_len carry P == 0 but their V rows are arbitrary pool / stale bytes: zero them.
      if (kv0 + kDecTile > kv_len) {
        const int st = j % NST;
        ptx::mbar_wait(bar(L::bVFull + st), static_cast<uint32_t>(j / NST) & 1u);
        if (kv0 + r >= kv_len) {
          uint4* row0 = reinterpret_cast<uint4*>(smem_gen + L::kV + st * kDecTileBytes + r * 128);
          uint4* row1 = reinterpret_cast<uint4*>(smem_gen + L::kV + st * kDecTileBytes + kDecHalf + r * 128);
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            row0[e] = z;
            row1[e] = z;
          }
        }
      }
      ptx::fence_proxy_async_smem();  // P^T (and zeroed V rows) visible to the tensor core's shared-memory reads
      ptx::tmem_wait_st();
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(bar(L::bPFull));
    }

    // ---- epilogue: this thread now owns head dim d = r of O^T --------------------------------------------------------
    float* red_l = red + (n_tiles & 1) * 64;  // the buffer the last tile did not use
#pragma unroll
    for (int h = 0; h < GP; ++h) {
      float l = l_thr[h];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
      if (lane == 0) red_l[warp * 16 + h] = l;
    }
    ptx::named_bar_sync(1, kDecTile);
    ptx::mbar_wait(bar(L::bOFull), 0);
    ptx::tc_fence_after_sync();
    uint32_t ov[16];
    ptx::tmem_ld_x16(tmem_o, ov);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int h = 0; h < GP; ++h) {
      if (h < a.group) {
        const float l = (red_l[h] + red_l[16 + h]) + (red_l[32 + h] + red_l[48 + h]);
        const int head = kvh * a.group + h;
        if (a.n_splits == 1) {
          T* orow = static_cast<T*>(a.out) + static_cast<int64_t>(t) * a.out_row_stride + head * kDecD;
          orow[r] = Elem<T>::from_f32(__uint_as_float(ov[h]) / l);
        } else {
          const int64_t pidx = (static_cast<int64_t>(t) * a.n_qo_heads + head) * a.n_splits + sp;
          a.part_o[pidx * kDecD + r] = __uint_as_float(ov[h]);
          if (r == 0) {
            a.part_ml[pidx * 2 + 0] = m_ref[h];
            a.part_ml[pidx * 2 + 1] = l;
          }
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kDecTmemCols);
  }
}

bool attn_decode_tc_supported(const HiAttnArgs& args) {
  const int group = args.n_kv_heads > 0 ? args.n_qo_heads / args.n_kv_heads : 0;
  return (args.dtype == HI_F16 || args.dtype == HI_BF16) && args.head_dim == kDecD && group >= 1 && group <= kDecN &&
         args.block_size >= 8 && args.block_size <= kDecTile && (kDecTile % args.block_size) == 0 &&
         (args.q_row_stride % 8) == 0 && aligned_to(args.q, 16) && aligned_to(args.key_cache, 16) &&
         aligned_to(args.value_cache, 16) && args.n_blocks > 0;
}

template <typename T, int GP>
static int launch_decode_t(const HiAttnArgs& args, const DecArgs& a, const CUtensorMap& mq, const CUtensorMap& mk,
                           const CUtensorMap& mv, cudaStream_t stream) {
  using L = DecSmem<1>;
  static PerDeviceFlags configured;
  HI_CUDA(configure_dynamic_smem(configured, paged_decode_tc_kernel<T, 1, GP>, L::kDynamicBytes));
  const dim3 grid(args.n_tokens, args.n_kv_heads, a.n_splits);
  timing_mark_start(stream);
  HI_CUDA(launch_pdl(paged_decode_tc_kernel<T, 1, GP>, grid, dim3(kDecThreads), L::kDynamicBytes, stream, mq, mk, mv, a));
  timing_mark_stop(stream);
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

int launch_attn_decode_tc(const HiAttnArgs& args, cudaStream_t stream) {
  if (!attn_decode_tc_supported(args)) {
    set_error("paged_attention: the tcgen05 decode path needs fp16/bf16, head_dim 128, group <= 16, block_size in {8,...,128}");
    return HI_ERR_UNSUPPORTED;
  }
  DecArgs a{};
  a.out = args.out;
  a.out_row_stride = args.out_row_stride;
  a.q_cu = args.q_cu_seq_lens;
  a.kv_cu = args.kv_cu_seq_lens;
  a.block_tables = args.block_tables;
  a.cu_blocks = args.cu_blocks_lens;
  a.n_seqs = args.n_seqs;
  a.n_qo_heads = args.n_qo_heads;
  a.n_kv_heads = args.n_kv_heads;
  a.group = args.n_qo_heads / args.n_kv_heads;
  a.block_size = args.block_size;
  a.scale_log2 = args.softmax_scale * 1.4426950408889634f;

  // split-KV.  With >= 1 CTA per SM and sequences of similar length the kernel already streams at the HBM roofline and
  // splitting only adds tail waves and merge traffic; ragged batches (longest sequence well above the mean) and launches
  // with fewer CTAs than SMs are cut into ~300 CTAs of equal work, never finer than 4 tiles (512 tokens) and never coarser
  // than 16 (larger launches run several waves and want the finer grain for balance).  Swept on B200 with the host out of
  // the picture (CUDA graph of 10 calls, profiles/r01_decode_tc_split_sweep.txt): 4 x 8k rows 16.8 us at 16 splits (22.0 at
  // 32, 86 unsplit), 16 ragged rows 32.2 us at 8 (36.9 at 16), the 48 ragged config-3 rows 69.3 us at 3 (75.5 at 6), 64 ragged
  // rows of the 72B shape 175.8 us at 4-6 (179 at 3).
  constexpr int kMinTilesPerSplit = 4, kMaxTilesPerSplit = 16;
  constexpr double kTargetCtas = 300.0;
  const int64_t base_ctas = static_cast<int64_t>(args.n_tokens) * args.n_kv_heads;
  const int max_tiles = (args.max_kv_len + kDecTile - 1) / kDecTile;
  // mean tiles per row from the block-table size (rows of one sequence share its blocks: exact for decode batches)
  double mean_tiles = static_cast<double>(max_tiles);
  if (args.kv_blocks_hint > 0 && args.n_seqs > 0)
    mean_tiles = static_cast<double>(args.kv_blocks_hint) * args.block_size / kDecTile / args.n_seqs;
  if (mean_tiles < 1.0) mean_tiles = 1.0;
  int n_splits = 1;
  if (base_ctas < 148 || max_tiles > 1.3 * mean_tiles) {
    const double total_work = mean_tiles * static_cast<double>(base_ctas);  // tile-CTA units
    int tps = static_cast<int>(total_work / kTargetCtas + 0.999);
    if (tps < kMinTilesPerSplit) tps = kMinTilesPerSplit;
    if (tps > kMaxTilesPerSplit) tps = kMaxTilesPerSplit;
    n_splits = (max_tiles + tps - 1) / tps;
  }
  const int max_splits = (max_tiles + kMinTilesPerSplit - 1) / kMinTilesPerSplit;
  if (n_splits > max_splits) n_splits = max_splits;
  if (const char* env = tuning_env("HI_DEC_SPLITS")) n_splits = atoi(env);
  if (n_splits < 1) n_splits = 1;
  n_splits = cap_splits(n_splits, args.n_tokens, args.n_qo_heads, kDecD);
  {
    const int64_t need = partial_bytes_per_split(args.n_tokens, args.n_qo_heads, kDecD) * n_splits;
    if (n_splits > 1 && (args.workspace == nullptr || need > args.workspace_bytes)) {
      set_error("paged_attention: workspace of %lld bytes is smaller than the %lld needed for %d KV splits (see hi_attention_workspace_bytes)",
                (long long)args.workspace_bytes, (long long)need, n_splits);
      return HI_ERR_WORKSPACE;
    }
  }
  a.tiles_per_split = (max_tiles + n_splits - 1) / n_splits;
  a.n_splits = (max_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  if (a.n_splits > 1) {
    const int64_t entries = static_cast<int64_t>(args.n_tokens) * args.n_qo_heads * a.n_splits;
    a.part_o = static_cast<float*>(args.workspace);
    a.part_ml = a.part_o + entries * kDecD;
  }

  CUtensorMap mq, mk, mv;
  int rc = make_map(&mq, args.dtype, args.q, args.n_tokens, args.n_qo_heads, args.q_row_stride, a.group, 1);
  if (rc != HI_OK) return rc;
  const int64_t n_slots = args.n_blocks * args.block_size;
  rc = pool_map(&mk, args.dtype, args.key_cache, n_slots, args.n_kv_heads, args.block_size);
  if (rc != HI_OK) return rc;
  rc = pool_map(&mv, args.dtype, args.value_cache, n_slots, args.n_kv_heads, args.block_size);
  if (rc != HI_OK) return rc;

  const bool wide = a.group > 8;
  if (args.dtype == HI_BF16) {
    rc = wide ? launch_decode_t<__nv_bfloat16, 16>(args, a, mq, mk, mv, stream) : launch_decode_t<__nv_bfloat16, 8>(args, a, mq, mk, mv, stream);
  } else {
    rc = wide ? launch_decode_t<__half, 16>(args, a, mq, mk, mv, stream) : launch_decode_t<__half, 8>(args, a, mq, mk, mv, stream);
  }
  if (rc != HI_OK || a.n_splits == 1) return rc;

  SimtArgs m{};
  m.out = args.out;
  m.out_row_stride = args.out_row_stride;
  m.q_cu = args.q_cu_seq_lens;
  m.kv_cu = args.kv_cu_seq_lens;
  m.n_seqs = args.n_seqs;
  m.n_tokens = args.n_tokens;
  m.n_qo_heads = args.n_qo_heads;
  m.n_chunks = a.n_splits;
  m.chunk_tiles = a.tiles_per_split * (kDecTile / 16);
  m.part_o = a.part_o;
  m.part_ml = a.part_ml;
  return launch_merge_partials(m, args.dtype, kDecD, stream);
}

}  // namespace hi
