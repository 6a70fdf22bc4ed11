// Paged causal attention on the 5th-gen tensor cores: tcgen05.mma with TMEM accumulators, operands staged by TMA.
//
// Same function as TorchCausalGroupedQueryPageAttentionHandler.forward (reference
// hydrainfer/layer/causal_attention.py:307-374) for q_len > 1 rows (prefill / chunked prefill) and, through
// GQA packing, for decode rows of grouped models.  It supersedes the reference's FA2 `flash_fwd_splitkv_kernel`
// (csrc/kernel/flash_attn/src/flash_fwd_launch_template.h:71-141: mma.sync m16n8k16, kBlockM = 64, one q-head per
// CTA) and flashinfer's BatchPrefillWithPagedKVCache (mma.sync as well).
//
// One CTA = one 128-row query tile of one (sequence, KV head):
//   * rows are (token, head-in-group) pairs, row = token * G + g, so the G query heads that share a KV head are
//     packed into the same MMA M dimension and every K/V byte staged in shared memory serves all of them;
//     the tile holds TQ = 128 / G tokens (one TMA box {64 dims, G heads, TQ tokens} per 64-dim half);
//   * KV is walked in 128-token tiles = 128/block_size pages; each page of a KV head is one TMA box
//     {64 dims, 1 head, block_size slots} per half of the head dim, written 128B-swizzled, so a tile is exactly
//     the K-major (for Q.K^T) / MN-major (for P.V) canonical UMMA layout with no data movement by threads;
//   * S = Q.K^T (M=128, N=128, K=128) accumulates in TMEM columns [0,128); the 4 softmax warps read their row
//     with tcgen05.ld (one thread = one row = one TMEM lane), apply the bottom-right causal mask
//     (key j visible to row i iff j <= i + L - q, causal_attention.py:339-342), keep a running max / sum, and
//     write P as packed 16-bit pairs back into TMEM columns [0,64) (aliasing S);
//   * O += P.V is issued with A = P from TMEM (tcgen05.mma ..., [tmem_a], ...) and B = V from shared memory;
//     O lives in TMEM columns [128,256) for the whole tile and is rescaled in place only when the running max
//     grows by more than 2^8 (lazy rescale), normalised by the row sum in the epilogue.
//   * warp roles: warps 0-3 softmax + epilogue, warp 4 TMA producer (and TMEM allocator), warp 5 MMA issuer;
//     all hand-offs are mbarriers (TMA transaction bytes, tcgen05.commit, thread arrivals).
// Two CTAs are co-resident per SM (256 TMEM columns and ~97 KiB of shared memory each) so one CTA's softmax
// overlaps the other's MMAs.
#include <cuda.h>

#include <cstdlib>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "attn_common.cuh"
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tma_maps.h"

namespace hi {

constexpr int kTcThreads = 192;
constexpr int kTileM = 128;
constexpr int kTileN = 128;
constexpr int kHeadDim = 128;
constexpr int kHalfBytes = kTileM * 128;       // one 64-dim half of a 128-row tile: 16 KiB
constexpr int kTileBytes = 2 * kHalfBytes;     // 32 KiB
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0;                  // S accumulator (fp32) and, aliased, P (16-bit pairs)
constexpr uint32_t kColO = 128;                // O accumulator (fp32)
constexpr float kRescaleThreshold = 8.0f;      // log2 of the largest stale-max overshoot P may carry

struct TcArgs {
  void* out;
  int64_t out_row_stride;  // elements
  const int32_t* q_cu;
  const int32_t* kv_cu;
  const int32_t* block_tables;
  const int32_t* cu_blocks;
  int n_qo_heads, n_kv_heads, group, block_size;
  int tq;            // query tokens per tile: 128 / group
  float scale_log2;  // softmax_scale * log2(e)
  int serialize;     // debug: wait for P.V to finish before the next Q.K^T is issued (HI_TC_SERIALIZE=1)
  // split-KV mode (n_splits > 1): CTA `sp` of a tile covers KV tiles [sp*tiles_per_split, (sp+1)*tiles_per_split) and
  // writes an fp32 partial (o, m, l) per row; merge_partials_kernel reduces them.
  int n_splits;
  int tiles_per_split;  // in 128-token tiles
  float* part_o;        // [n_tokens * Hq * n_splits][128]
  float* part_ml;       // [n_tokens * Hq * n_splits][2]
};

// D = head_dim (128 or 256): a 128-row Q tile and a 128-key K or V tile are D / 64 halves of 16 KiB each.
template <int NST, int D = kHeadDim>
struct TcSmem {
  static constexpr int kTile = (D / 64) * kHalfBytes;   // 32 KiB at head_dim 128, 64 KiB at 256
  static constexpr int kQ = 0;
  static constexpr int kK = kTile;
  static constexpr int kV = kTile + NST * kTile;
  static constexpr int kBars = kTile + 2 * NST * kTile;
  // barrier slots (8 bytes each)
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
  static constexpr int kDynamicBytes = kTotal + 1024;  // slack to align the base to 1024 B (128B swizzle atoms)
};

// D = 256 (round 2): the reference's fused backend covers head_dim 256 (static_switch.h:70-85).  One 128-row tile per CTA with
// S (128 TMEM columns) | O (256 columns); Q, one K and one V stage are 64 KiB each, so one CTA per SM and a single stage:
// the serial chain Q.K^T -> softmax -> P.V of this kernel, at twice the tensor work per step.
template <typename T, int NST, int D = kHeadDim>
__global__ void __launch_bounds__(kTcThreads, (NST == 1 && D == kHeadDim) ? 2 : 1)
paged_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const TcArgs a) {
  using L = TcSmem<NST, D>;
  constexpr bool kBf16 = sizeof(T) == 2 && !std::is_same<T, __half>::value;
  constexpr int kHalves = D / 64;                         // 64-dim halves of a row
  constexpr uint32_t kTmemAlloc = D == kHeadDim ? kTmemCols : 512u;  // S 128 + O D columns, rounded up to a power of two

  // ---- which tile --------------------------------------------------------------------------------------------------
  const int b = blockIdx.z;
  const int kvh = blockIdx.y;
  const int q_start = __ldg(a.q_cu + b);
  const int q_len = __ldg(a.q_cu + b + 1) - q_start;
  const int kv_len = __ldg(a.kv_cu + b + 1) - __ldg(a.kv_cu + b);
  const int n_q_tiles = (q_len + a.tq - 1) / a.tq;
  // Heaviest tiles (the end of the sequence sees the most keys) are scheduled first.
  const int sp = static_cast<int>(blockIdx.x) % a.n_splits;
  const int q_tile = n_q_tiles - 1 - static_cast<int>(blockIdx.x) / a.n_splits;
  if (q_tile < 0) return;
  const int i0 = q_tile * a.tq;                                   // first query position of the tile
  const int i_last = min(q_len, i0 + a.tq) - 1;                   // last valid query position
  const int kv_end = i_last + (kv_len - q_len) + 1;               // keys [0, kv_end) are visible to the tile
  const int n_kv_tiles_all = (kv_end + kTileN - 1) / kTileN;
  const int j_begin = sp * a.tiles_per_split;
  if (j_begin >= n_kv_tiles_all) return;  // this split has no visible key for the tile; the merge skips it too
  const int n_kv_tiles = min(n_kv_tiles_all - j_begin, a.tiles_per_split);  // tiles this CTA walks
  const int blk0 = __ldg(a.cu_blocks + b);
  const int n_pages = __ldg(a.cu_blocks + b + 1) - blk0;
  const int pages_per_tile = kTileN / a.block_size;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  auto bar = [&](int idx) -> uint32_t { return smem_base + L::kBars + idx * 8; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kTmemPtr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup ------------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar(L::bQFull), 1);
    for (int s = 0; s < NST; ++s) {
      ptx::mbar_init(bar(L::bKFull + s), 1);
      ptx::mbar_init(bar(L::bKEmpty + s), 1);
      ptx::mbar_init(bar(L::bVFull + s), 1);
      ptx::mbar_init(bar(L::bVEmpty + s), 1);
    }
    ptx::mbar_init(bar(L::bSFull), 1);
    ptx::mbar_init(bar(L::bPFull), kTileM);  // every softmax thread arrives
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
    ptx::tmem_alloc(smem_base + L::kTmemPtr, kTmemAlloc);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 4) {
    // ================================================ TMA producer ================================================
    if (lane == 0) {
      // Q tile: both 64-dim halves, rows ordered (token, head-in-group).
      const uint32_t q_bytes = static_cast<uint32_t>(kHalves) * static_cast<uint32_t>(a.group * a.tq) * 128u;
      ptx::mbar_arrive_expect_tx(bar(L::bQFull), q_bytes);
#pragma unroll
      for (int h = 0; h < kHalves; ++h)
        ptx::tma_load_3d(smem_base + L::kQ + h * kHalfBytes, &tm_q, bar(L::bQFull), 64 * h, kvh * a.group, q_start + i0);
    }
    const uint32_t page_half_bytes = static_cast<uint32_t>(a.block_size) * 128u;
    for (int j = 0; j < n_kv_tiles; ++j) {
      const int st = j % NST;
      const uint32_t ph = static_cast<uint32_t>(j / NST) & 1u;
      const int page0 = (j_begin + j) * pages_per_tile;
      const int n_valid = min(pages_per_tile, n_pages - page0);
      // lane p stages page p of the tile
      int blk = 0;
      if (lane < n_valid) blk = __ldg(a.block_tables + blk0 + page0 + lane);
      const uint32_t tx = static_cast<uint32_t>(n_valid) * static_cast<uint32_t>(kHalves) * page_half_bytes;
      // K
      if (lane == 0) {
        ptx::mbar_wait(bar(L::bKEmpty + st), ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar(L::bKFull + st), tx);
      }
      __syncwarp();
      if (lane < n_valid) {
        const uint32_t dst = smem_base + L::kK + st * L::kTile + lane * page_half_bytes;
#pragma unroll
        for (int h = 0; h < kHalves; ++h) ptx::tma_load_3d(dst + h * kHalfBytes, &tm_k, bar(L::bKFull + st), 64 * h, kvh, blk * a.block_size);
      }
      // V
      if (lane == 0) {
        ptx::mbar_wait(bar(L::bVEmpty + st), ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar(L::bVFull + st), tx);
      }
      __syncwarp();
      if (lane < n_valid) {
        const uint32_t dst = smem_base + L::kV + st * L::kTile + lane * page_half_bytes;
#pragma unroll
        for (int h = 0; h < kHalves; ++h) ptx::tma_load_3d(dst + h * kHalfBytes, &tm_v, bar(L::bVFull + st), 64 * h, kvh, blk * a.block_size);
      }
    }
  } else if (warp == 5) {
    // ================================================ MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_f16(kBf16, false, false, kTileM, kTileN);
      constexpr uint32_t idesc_pv = ptx::make_idesc_f16(kBf16, false, true, kTileM, D);
      const uint32_t tmem_s = tmem_base + kColS;
      const uint32_t tmem_o = tmem_base + kColO;
      ptx::mbar_wait(bar(L::bQFull), 0);
      for (int j = 0; j < n_kv_tiles; ++j) {
        const int st = j % NST;
        const uint32_t ph = static_cast<uint32_t>(j / NST) & 1u;
        // ---- S = Q . K^T : 2 halves x 4 k-steps of 16 dims; both operands K-major, 8-row groups 1024 B apart
        ptx::mbar_wait(bar(L::bKFull + st), ph);
        ptx::tc_fence_after_sync();
        const uint32_t q_addr = smem_base + L::kQ;
        const uint32_t k_addr = smem_base + L::kK + st * L::kTile;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {  // D / 64 halves x 4 k-steps of 16 dims
          const uint32_t off = (kk >> 2) * kHalfBytes + (kk & 3) * 32;
          ptx::mma_f16_ss(tmem_s, ptx::make_smem_desc_sw128(q_addr + off, 16, 1024),
                          ptx::make_smem_desc_sw128(k_addr + off, 16, 1024), idesc_qk, kk > 0);
        }
        ptx::mma_commit(bar(L::bKEmpty + st));  // K slot reusable once these MMAs have read it
        ptx::mma_commit(bar(L::bSFull));        // S ready (also implies the previous P.V finished)
        // ---- O += P . V : 8 k-steps of 16 tokens; A = P in TMEM (8 columns per step), B = V MN-major
        ptx::mbar_wait(bar(L::bPFull), static_cast<uint32_t>(j) & 1u);
        ptx::mbar_wait(bar(L::bVFull + st), ph);
        ptx::tc_fence_after_sync();
        const uint32_t v_addr = smem_base + L::kV + st * L::kTile;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // (N = D: the MN-major descriptor walks the D / 64 halves at its leading-dim offset)
          ptx::mma_f16_ts(tmem_o, tmem_s + kk * 8, ptx::make_smem_desc_sw128(v_addr + kk * 2048, kHalfBytes, 1024),
                          idesc_pv, (j > 0) || (kk > 0));
        }
        ptx::mma_commit(bar(L::bVEmpty + st));
        if (j == n_kv_tiles - 1) ptx::mma_commit(bar(L::bOFull));
        if (a.serialize) ptx::mbar_wait(bar(L::bVEmpty + st), ph);
      }
    }
  } else {
    // ================================================ softmax + epilogue ===========================================
    const int r = threadIdx.x;                       // tile row == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + kColS;
    const uint32_t tmem_o = tmem_base + lane_base + kColO;
    const int tok = r / a.group;                     // token within the tile
    const int g = r - tok * a.group;
    const int i = i0 + tok;                          // query position within the sequence
    const bool row_valid = (tok < a.tq) && (i < q_len);
    const int lim = i + (kv_len - q_len);            // last visible key index for this row
    float m_used = 0.f;                              // exponent reference, scaled log2 domain
    float l = 0.f;
    // Rows past the tile's valid (token, head) pairs are padding: a warp that owns only padding rows skips the softmax
    // math (decode tiles of grouped models have <= 16 real rows).  Its rows of P keep stale bits; MMA rows are independent.
    const int rows_real = (min(q_len, i0 + a.tq) - i0) * a.group;
    const bool warp_active = warp * 32 < rows_real;

    for (int j = 0; j < n_kv_tiles; ++j) {
      const int kv0 = (j_begin + j) * kTileN;
      const int col_lim = lim - kv0;                 // columns [0, col_lim] are visible
      const bool need_mask = col_lim < kTileN - 1;
      ptx::mbar_wait(bar(L::bSFull), static_cast<uint32_t>(j) & 1u);
      ptx::tc_fence_after_sync();
      if (warp_active) {

      // Two sweeps over the row's 128 scores, 64 columns per TMEM round trip.  The mask is only evaluated on tiles that
      // touch the causal diagonal or the end of the sequence, and the arithmetic uses the 3-input max and the packed
      // fp32x2 FMA/ADD of sm_100, so that MUFU.EX2 (16 lanes/clk/SM) is the only saturated pipe.
      const bool mask_tile = __any_sync(0xffffffffu, need_mask);
      uint32_t va[32], vb[32];
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ptx::tmem_ld_x32(tmem_s + h * 64, va);
        ptx::tmem_ld_x32(tmem_s + h * 64 + 32, vb);
        ptx::tmem_wait_ld();
        if (mask_tile) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (h * 64 + e > col_lim) va[e] = 0xff800000u;  // -inf
            if (h * 64 + 32 + e > col_lim) vb[e] = 0xff800000u;
          }
        }
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            mx4[u] = fmax3(mx4[u], __uint_as_float(va[e + 2 * u]), __uint_as_float(va[e + 2 * u + 1]));
            mx4[u] = fmax3(mx4[u], __uint_as_float(vb[e + 2 * u]), __uint_as_float(vb[e + 2 * u + 1]));
          }
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float mxs = mx * a.scale_log2;
      if (j == 0) {
        m_used = (mxs == -INFINITY) ? 0.f : mxs;
      } else if (__any_sync(0xffffffffu, r < rows_real && mxs > m_used + kRescaleThreshold)) {  // padding rows (stale Q) do not vote
        // Lazy rescale: the whole warp pays the TMEM round trip only when some row's max grew by > 2^8.
        const float m_new = fmaxf(m_used, mxs);
        const float alpha = fast_exp2(m_used - m_new);
        l *= alpha;
        m_used = m_new;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          ptx::tmem_ld_x32(tmem_o + c * 32, va);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) va[e] = __float_as_uint(__uint_as_float(va[e]) * alpha);
          ptx::tmem_st_x32(tmem_o + c * 32, va);
        }
      }

      // P = exp2(S * scale - m) as 16-bit pairs, written over S (P columns [16c, 16c+16) only cover S columns already read)
      const float2 sc2 = make_float2(a.scale_log2, a.scale_log2);
      const float2 nm2 = make_float2(-m_used, -m_used);
      float2 ls2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      auto exp_pack_store = [&](uint32_t (&v)[32], int c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float s0 = __uint_as_float(v[e]), s1 = __uint_as_float(v[e + 1]);
          if (mask_tile) {
            if (c * 32 + e > col_lim) s0 = -INFINITY;
            if (c * 32 + e + 1 > col_lim) s1 = -INFINITY;
          }
          const float2 t2 = ffma2(make_float2(s0, s1), sc2, nm2);
          const float2 p2 = make_float2(fast_exp2(t2.x), fast_exp2(t2.y));
          ls2[(e >> 1) & 1] = fadd2(ls2[(e >> 1) & 1], p2);
          pk[e >> 1] = pack2<T>(p2.x, p2.y);
        }
        ptx::tmem_st_x16(tmem_s + c * 16, pk);
      };
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ptx::tmem_ld_x32(tmem_s + h * 64, va);
        ptx::tmem_ld_x32(tmem_s + h * 64 + 32, vb);
        ptx::tmem_wait_ld();
        exp_pack_store(va, 2 * h);
        exp_pack_store(vb, 2 * h + 1);
      }
      const float lsum = (ls2[0].x + ls2[0].y) + (ls2[1].x + ls2[1].y);
      l += lsum;
      }  // warp_active

      // Keys at or beyond kv_len (tail of the last page, pages that do not exist) carry P == 0, but their V rows are
      // whatever the pool / stale shared memory holds; zero them so 0 * NaN cannot reach O.
      if (kv0 + kTileN > kv_len) {
        const int st = j % NST;
        ptx::mbar_wait(bar(L::bVFull + st), static_cast<uint32_t>(j / NST) & 1u);
        if (kv0 + r >= kv_len) {
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int h = 0; h < kHalves; ++h) {
            uint4* row = reinterpret_cast<uint4*>(smem_gen + L::kV + st * L::kTile + h * kHalfBytes + r * 128);
#pragma unroll
            for (int e = 0; e < 8; ++e) row[e] = z;
          }
        }
        ptx::fence_proxy_async_smem();
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(bar(L::bPFull));
    }

    // ---- epilogue: O / l -> out ----------------------------------------------------------------------------------
    ptx::mbar_wait(bar(L::bOFull), 0);
    ptx::tc_fence_after_sync();
    const float inv_l = 1.f / l;
    T* orow = static_cast<T*>(a.out) + static_cast<int64_t>(q_start + i) * a.out_row_stride +
              (kvh * a.group + g) * D;
    const int64_t pidx = (static_cast<int64_t>(q_start + i) * a.n_qo_heads + (kvh * a.group + g)) * a.n_splits + sp;
    if (a.n_splits > 1 && row_valid) {
      a.part_ml[pidx * 2 + 0] = m_used;
      a.part_ml[pidx * 2 + 1] = l;
    }
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t v[32];
      ptx::tmem_ld_x32(tmem_o + c * 32, v);
      ptx::tmem_wait_ld();
      if (a.n_splits > 1) {
        if (row_valid) {
          float4* dst = reinterpret_cast<float4*>(a.part_o + pidx * D + c * 32);
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            dst[e >> 2] = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        }
      } else if (row_valid) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 w;
          w.x = pack2<T>(__uint_as_float(v[e + 0]) * inv_l, __uint_as_float(v[e + 1]) * inv_l);
          w.y = pack2<T>(__uint_as_float(v[e + 2]) * inv_l, __uint_as_float(v[e + 3]) * inv_l);
          w.z = pack2<T>(__uint_as_float(v[e + 4]) * inv_l, __uint_as_float(v[e + 5]) * inv_l);
          w.w = pack2<T>(__uint_as_float(v[e + 6]) * inv_l, __uint_as_float(v[e + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + e) = w;
        }
      }
    }
  }

  // ---- teardown ----------------------------------------------------------------------------------------------------
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemAlloc);
  }
}

// ---- host: tensor maps ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (cached == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HI_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return HI_ERR_CUDA;
    }
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return HI_OK;
}

// 3-D map over [rows, heads, 128] 16-bit elements; box = {64 dims, box_heads, box_rows}, 128B swizzle.
int make_map(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t heads, int64_t row_stride_elems,
                    int box_heads, int box_rows) {
  return make_map_d(map, dtype, base, rows, heads, kHeadDim, row_stride_elems, box_heads, box_rows);
}

int make_map_d(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t heads, int head_dim, int64_t row_stride_elems,
               int box_heads, int box_rows) {
  EncodeTiledFn encode = nullptr;
  const int rc = get_encode_fn(&encode);
  if (rc != HI_OK) return rc;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(head_dim), static_cast<cuuint64_t>(heads), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(head_dim) * 2, static_cast<cuuint64_t>(row_stride_elems) * 2};
  const cuuint32_t box[3] = {64u, static_cast<cuuint32_t>(box_heads), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult res = encode(map, dtype == HI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                              const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld heads=%lld stride=%lld box=%d,%d)", (int)res,
              (long long)rows, (long long)heads, (long long)row_stride_elems, box_heads, box_rows);
    return HI_ERR_CUDA;
  }
  return HI_OK;
}

// Pool maps are reused by every layer call with the same base pointer; cache them.
struct MapKey {
  const void* base;
  int64_t rows, heads;
  int dtype, box_rows;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && heads == o.heads && dtype == o.dtype && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.base) ^ (std::hash<int64_t>()(k.rows) * 31) ^ (std::hash<int64_t>()(k.heads) * 131) ^
           (static_cast<size_t>(k.dtype) << 8) ^ (static_cast<size_t>(k.box_rows) << 16);
  }
};
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_pool_maps;

int pool_map_d(CUtensorMap* out, int dtype, const void* base, int64_t n_slots, int heads, int head_dim, int block_size) {
  const MapKey key{base, n_slots, heads, dtype | (head_dim << 8), block_size};
  std::lock_guard<std::mutex> lock(g_map_mu);
  auto it = g_pool_maps.find(key);
  if (it != g_pool_maps.end()) {
    *out = it->second;
    return HI_OK;
  }
  CUtensorMap m;
  const int rc = make_map_d(&m, dtype, base, n_slots, heads, head_dim, static_cast<int64_t>(heads) * head_dim, 1, block_size);
  if (rc != HI_OK) return rc;
  if (g_pool_maps.size() > 4096) g_pool_maps.clear();
  g_pool_maps.emplace(key, m);
  *out = m;
  return HI_OK;
}

int pool_map(CUtensorMap* out, int dtype, const void* base, int64_t n_slots, int heads, int block_size) {
  return pool_map_d(out, dtype, base, n_slots, heads, kHeadDim, block_size);
}

bool attn_tc_supported(const HiAttnArgs& args) {
  const int group = args.n_kv_heads > 0 ? args.n_qo_heads / args.n_kv_heads : 0;
  return (args.dtype == HI_F16 || args.dtype == HI_BF16) && (args.head_dim == kHeadDim || args.head_dim == 256) && group >= 1 && group <= kTileM &&
         args.block_size >= 8 && args.block_size <= kTileN && (kTileN % args.block_size) == 0 &&
         (args.q_row_stride % 8) == 0 && (args.out_row_stride % 8) == 0 && aligned_to(args.q, 16) &&
         aligned_to(args.out, 16) && aligned_to(args.key_cache, 16) && aligned_to(args.value_cache, 16) &&
         args.n_blocks > 0;
}

template <typename T, int NST, int D = kHeadDim>
static int launch_tc_t(const HiAttnArgs& args, const TcArgs& a, const CUtensorMap& mq, const CUtensorMap& mk,
                       const CUtensorMap& mv, cudaStream_t stream) {
  using L = TcSmem<NST, D>;
  static PerDeviceFlags configured;
  HI_CUDA(configure_dynamic_smem(configured, paged_attn_tc_kernel<T, NST, D>, L::kDynamicBytes));
  const int q_tiles = (args.max_q_len + a.tq - 1) / a.tq;
  const dim3 grid(q_tiles * a.n_splits, args.n_kv_heads, args.n_seqs);
  timing_mark_start(stream);
  paged_attn_tc_kernel<T, NST, D><<<grid, kTcThreads, L::kDynamicBytes, stream>>>(mq, mk, mv, a);
  timing_mark_stop(stream);
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

int launch_attn_tc(const HiAttnArgs& args, cudaStream_t stream) {
  if (!attn_tc_supported(args)) {
    set_error("paged_attention: the tcgen05 path needs fp16/bf16, head_dim 128 or 256, block_size in {8,16,32,64,128} and 16-byte aligned rows");
    return HI_ERR_UNSUPPORTED;
  }
  TcArgs a{};
  a.out = args.out;
  a.out_row_stride = args.out_row_stride;
  a.q_cu = args.q_cu_seq_lens;
  a.kv_cu = args.kv_cu_seq_lens;
  a.block_tables = args.block_tables;
  a.cu_blocks = args.cu_blocks_lens;
  a.n_qo_heads = args.n_qo_heads;
  a.n_kv_heads = args.n_kv_heads;
  a.group = args.n_qo_heads / args.n_kv_heads;
  a.block_size = args.block_size;
  a.tq = kTileM / a.group;
  a.scale_log2 = args.softmax_scale * 1.4426950408889634f;
  {
    const char* env = tuning_env("HI_TC_SERIALIZE");
    a.serialize = (env != nullptr && env[0] == '1') ? 1 : 0;
  }

  // ---- split-KV: only for launches too small to fill the machine (decode batches, a few long prompts) ---------------
  constexpr int kSplitTargetCtas = 592;  // 148 SMs x 4
  constexpr int kMinTilesPerSplit = 2;   // never finer than 256 tokens
  const int q_tiles = (args.max_q_len + a.tq - 1) / a.tq;
  const int64_t base_ctas = static_cast<int64_t>(q_tiles) * args.n_kv_heads * args.n_seqs;
  const int max_kv_tiles = (args.max_kv_len + kTileN - 1) / kTileN;
  int n_splits = 1;
  if (base_ctas < kSplitTargetCtas) {
    n_splits = static_cast<int>((kSplitTargetCtas + base_ctas - 1) / base_ctas);
    const int max_splits = (max_kv_tiles + kMinTilesPerSplit - 1) / kMinTilesPerSplit;
    if (n_splits > max_splits) n_splits = max_splits;
    if (const char* env = tuning_env("HI_TC_SPLITS")) n_splits = atoi(env);  // tuning override
    if (n_splits < 1) n_splits = 1;
    n_splits = cap_splits(n_splits, args.n_tokens, args.n_qo_heads, args.head_dim);
    const int64_t need = partial_bytes_per_split(args.n_tokens, args.n_qo_heads, args.head_dim) * n_splits;
    if (n_splits > 1 && (args.workspace == nullptr || need > args.workspace_bytes)) {
      set_error("paged_attention: workspace of %lld bytes is smaller than the %lld needed for %d KV splits (see hi_attention_workspace_bytes)",
                (long long)args.workspace_bytes, (long long)need, n_splits);
      return HI_ERR_WORKSPACE;
    }
  }
  a.tiles_per_split = (max_kv_tiles + n_splits - 1) / n_splits;
  a.n_splits = (max_kv_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  if (a.n_splits > 1) {
    const int64_t entries = static_cast<int64_t>(args.n_tokens) * args.n_qo_heads * a.n_splits;
    a.part_o = static_cast<float*>(args.workspace);
    a.part_ml = a.part_o + entries * args.head_dim;
  }

  CUtensorMap mq, mk, mv;
  int rc = make_map_d(&mq, args.dtype, args.q, args.n_tokens, args.n_qo_heads, args.head_dim, args.q_row_stride, a.group, a.tq);
  if (rc != HI_OK) return rc;
  const int64_t n_slots = args.n_blocks * args.block_size;
  rc = pool_map_d(&mk, args.dtype, args.key_cache, n_slots, args.n_kv_heads, args.head_dim, args.block_size);
  if (rc != HI_OK) return rc;
  rc = pool_map_d(&mv, args.dtype, args.value_cache, n_slots, args.n_kv_heads, args.head_dim, args.block_size);
  if (rc != HI_OK) return rc;

  if (args.head_dim == 256) {  // 64 KiB each for Q, the K stage and the V stage: one stage, one CTA per SM
    rc = args.dtype == HI_BF16 ? launch_tc_t<__nv_bfloat16, 1, 256>(args, a, mq, mk, mv, stream)
                               : launch_tc_t<__half, 1, 256>(args, a, mq, mk, mv, stream);
    if (rc != HI_OK || a.n_splits == 1) return rc;
  } else {
  // Ring depth 1 = two co-resident CTAs per SM whose QK / softmax / PV phases overlap each other; measured better than
  // one CTA per SM with a 2- or 3-deep ring for both prefill and decode tiles (profiles/r01_notes.md).
  int stages = 1;
  if (const char* env = tuning_env("HI_TC_STAGES")) stages = atoi(env);
  if (args.dtype == HI_BF16) {
    rc = stages == 3 ? launch_tc_t<__nv_bfloat16, 3>(args, a, mq, mk, mv, stream)
       : stages == 2 ? launch_tc_t<__nv_bfloat16, 2>(args, a, mq, mk, mv, stream)
                     : launch_tc_t<__nv_bfloat16, 1>(args, a, mq, mk, mv, stream);
  } else {
    rc = stages == 3 ? launch_tc_t<__half, 3>(args, a, mq, mk, mv, stream)
       : stages == 2 ? launch_tc_t<__half, 2>(args, a, mq, mk, mv, stream)
                     : launch_tc_t<__half, 1>(args, a, mq, mk, mv, stream);
  }
  if (rc != HI_OK || a.n_splits == 1) return rc;
  }

  SimtArgs m{};
  m.out = args.out;
  m.out_row_stride = args.out_row_stride;
  m.q_cu = args.q_cu_seq_lens;
  m.kv_cu = args.kv_cu_seq_lens;
  m.n_seqs = args.n_seqs;
  m.n_tokens = args.n_tokens;
  m.n_qo_heads = args.n_qo_heads;
  m.n_chunks = a.n_splits;
  m.chunk_tiles = a.tiles_per_split * (kTileN / 16);
  m.part_o = a.part_o;
  m.part_ml = a.part_ml;
  return launch_merge_partials(m, args.dtype, args.head_dim, stream);
}

}  // namespace hi
