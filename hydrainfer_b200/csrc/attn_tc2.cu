// Paged causal attention on tcgen05 — the prefill / chunked-prefill kernel: two query tiles per CTA, S double-buffered.
//
// Same function as TorchCausalGroupedQueryPageAttentionHandler.forward (reference
// hydrainfer/layer/causal_attention.py:307-374) and the same operand staging as attn_tc.cu (TMA boxes per page written
// 128B-swizzled = canonical UMMA layouts, S and O in TMEM, P fed to the second product straight from TMEM).  What changes
// is the schedule, built around the two facts that bound attention on this part (B300_MICROARCH.md, tcgen05 / pipe rates):
// a 128-row x 64-key score tile costs 512 tensor-pipe cycles (Q.K^T + P.V) and 512 MUFU cycles (16 ex2/clk/SM), and
// every staged K/V byte costs L2 -> shared-memory bandwidth.
//
//   * One CTA owns TWO 128-row query tiles of the same (sequence, KV head) and ONE K/V ring: every staged K/V byte feeds
//     both tiles.  One CTA per SM; 512 TMEM columns = per tile { S[0] (64) | S[1] (64) | O (128) }.
//   * KV is walked in 64-key steps and S is DOUBLE-BUFFERED: the MMA lane of tile t issues
//         ... P_t.V(j)  Q_t.K(j+2)  P_t.V(j+1)  Q_t.K(j+3) ...
//     so S_t(j+1) is already complete when the softmax warpgroup of tile t finishes P_t(j): the softmax warps never wait for
//     the tensor pipe, and the tensor pipe only waits for P.  (With a single S buffer the chain
//     softmax -> P.V -> Q.K -> softmax is serial per tile and both pipes idle ~45 % of the time: measured.)
//   * Issue bandwidth is a first-class resource: a tcgen05.mma of these shapes runs in 32-64 cycles (tools/probes/
//     mma_probe.cu), and every extra instruction the issuing lane executes between two MMAs costs ~10 cycles.  So each tile
//     has its OWN MMA-issuing warp (9 and 10), K and V have their own TMA warps (8 and 11), and a step costs each MMA warp
//     3 waits + 3 commits; those warps stay converged and elect the issuing lane so that descriptors live in uniform
//     registers.
//   * Two softmax warpgroups (thread = row = TMEM lane): a thread reads its 64 scores from TMEM once, keeps them in
//     registers for max / exp2 / row sum / 16-bit pack, and writes P over the S buffer it came from.
//   * O is rescaled lazily (only when a row max grows by more than 2^8) by the softmax thread itself, after waiting for
//     P_t.V(j-1) (its own commit barrier; S_t(j) no longer implies it).
//   * The TMA producer zeroes V rows at or beyond kv_len of the last step (P is 0 there, the pool bytes are arbitrary).
//   * PERSISTENT: one CTA per SM walks the launch's work items (item = sequence x KV head x pair of tiles x split) with a
//     stride of gridDim.x.  TMEM, barriers and tensor maps are set up once, and the item boundary is pipelined like any
//     other step: ring stages, S/P buffers and barrier phases run on counters that never reset, the producers load the
//     next item's Q/K/V while the softmax warps are still in the epilogue of the current one (Q_t and O_t each have an
//     "empty" barrier), so short sequences do not pay a CTA launch + TMEM allocation + pipeline fill per tile.
// Rows are (token, head-in-group) pairs, so GQA groups share the staged K/V exactly as in attn_tc.cu.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "attn_common.cuh"
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tma_maps.h"

namespace hi {

constexpr int kP2Threads = 384;               // warps 0-3 / 4-7: softmax of tile 0 / 1; 8: K TMA; 9 / 10: MMA of tile 0 / 1; 11: V TMA
constexpr int kP2TileM = 128;
constexpr int kP2TileN = 64;                  // keys per step
constexpr int kP2D = 128;
constexpr int kP2QHalf = kP2TileM * 128;      // one 64-dim half of a 128-row Q tile: 16 KiB
constexpr int kP2QTile = 2 * kP2QHalf;        // 32 KiB
constexpr int kP2Half = kP2TileN * 128;       // one 64-dim half of a 64-key K or V step: 8 KiB
constexpr int kP2Tile = 2 * kP2Half;          // 16 KiB
constexpr uint32_t kP2TmemCols = 512;
constexpr uint32_t kP2ColTile = 256;          // columns per query tile: S[0] at +0, S[1] at +64, O at +128
constexpr uint32_t kP2ColO = 128;
constexpr float kP2Rescale = 8.0f;
constexpr int kP2Xchg = kP2D + 4;             // floats per row of the tile-1 -> tile-0 state exchange of an interleaved decode item: O[128], m, l, pad
constexpr int kP2ClaimAhead = 8;              // 64-key steps of the current item left (for the K producer) when the next item is claimed;
                                              // swept on B200: 4, 8 and 16 are equivalent, 32 loses 2.5 % on the config-3 mixed batch, whole-item
                                              // lookahead 26 %

struct P2Args {
  void* out;
  int64_t out_row_stride;
  const int32_t* q_cu;
  const int32_t* kv_cu;
  const int32_t* block_tables;
  const int32_t* cu_blocks;
  const int32_t* work_items;  // optional host plan: [n][2] = (sequence, pair of tiles), heaviest first; grid.x walks it
  int n_qo_heads, n_kv_heads, group, block_size;
  int tq;            // query tokens per 128-row tile: 128 / group
  float scale_log2;
  int n_splits;
  int tiles_per_split;  // in 64-key steps
  int direct_limit;     // a tile that sees at most this many steps (>= tiles_per_split) is ONE work item - split 0 walks all of it and
                        // writes `out` directly, its other splits are empty - so that only the heaviest tiles of a launch are cut
  float* part_o;
  float* part_ml;
  unsigned int* work_counter;  // zeroed before the launch; NULL = static (boustrophedon) assignment
  int n_items;       // work items of the launch = CTA-sized units: (sequence, KV head, pair of tiles, split)
  int max_pairs;     // without a host plan: pair slots per sequence, ceil(max_q_len / (2 tq))
  int interleave_decode;  // decode items (q_len == 1) are walked by BOTH tiles: tile t takes the 64-key steps of parity t (see P2Item::il)
  int use_tma_store;  // epilogue of whole direct tiles through shared memory + cp.async.bulk.tensor (needs tm_o and head_dim 128)
  int debug;  // timing experiments only (HI_PAIR_DEBUG): bit 0 = softmax warps skip their math, bit 1 = no MMA is issued,
              // bit 2 = K/V tiles are not loaded (barriers only)
  // ---- un-paged varlen mode (template parameter VL; hi_varlen_attention) --------------------------------------------------
  // K and V are plain [n_k_tokens, n_kv_heads, head_dim] tensors: key j of sequence b is row kv_cu[b] + j, a "page" is one
  // 64-key step, block_tables / cu_blocks are unused.  head_dim may be any multiple of 8 up to 128: the tensor maps carry the
  // real head_dim, TMA zero-fills the dims beyond it, Q.K^T walks ceil(head_dim / 16) k-steps and P.V produces
  // round_up(head_dim, 16) columns.
  int head_dim;
  int causal;        // 0: every key of the sequence is visible to every query row (vision encoders)
  int n_halves;      // 64-dim halves of a row that exist: 1 (head_dim <= 64) or 2
  int n_kk;          // 16-dim k-steps of Q.K^T
  uint32_t idesc_pv; // instruction descriptor of P.V with N = round_up(head_dim, 16)
};

template <int NK, int NV>
struct P2Smem {
  static constexpr int kQ = 0;                          // two tiles
  static constexpr int kK = 2 * kP2QTile;
  static constexpr int kV = kK + NK * kP2Tile;
  static constexpr int kOst = kV + NV * kP2Tile;        // output staging for the TMA-store epilogue: one 64-dim half (128 rows x 128 B) per tile
#ifdef HI_P2_NO_STAGING  // dev variant: the round-1 shared-memory footprint (192 KiB), per-row stores only
  static constexpr int kBars = kOst;
#else
  static constexpr int kBars = kOst + 2 * kP2QHalf;
#endif
  static constexpr int bQFull = 0;                      // [2]
  static constexpr int bKFull = 2;                      // [NK]
  static constexpr int bKEmpty = bKFull + NK;
  static constexpr int bVFull = bKEmpty + NK;           // [NV]
  static constexpr int bVEmpty = bVFull + NV;
  static constexpr int bSFull = bVEmpty + NV;           // [tile][buffer]
  static constexpr int bPFull = bSFull + 4;             // [tile][buffer]: a warpgroup may run two steps ahead of the MMA lane
  static constexpr int bPvDone = bPFull + 4;            // [2] P_t.V(n_t - 2) complete (lazy rescale in the last step)
  static constexpr int bOFull = bPvDone + 2;            // [2]
  static constexpr int bVTail = bOFull + 2;
  static constexpr int bQEmpty = bVTail + 1;            // [2] every Q_t.K of the item has read Q_t (next item's Q may land)
  static constexpr int bOEmpty = bQEmpty + 2;           // [2] the epilogue has read O_t out of TMEM (next item may overwrite it)
  static constexpr int bItemFull = bOEmpty + 2;         // [4] ring of work-item indices published by warp 8
  static constexpr int bItemEmpty = bItemFull + 4;      // [4] read by warp 11, both MMA warps (one lane each) and all 256 softmax threads
  static constexpr int kNumBars = bItemEmpty + 4;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kItemRing = kTmemPtr + 16;       // int[4]
  static constexpr int kTotal = kItemRing + 16;
  static constexpr int kDynamicBytes = kTotal + 1024;   // slack to align the base to 1024 B (128B-swizzle atoms)
};

// One work item, decoded identically by every role.
struct P2Item {
  int b, kvh, sp;
  int q_start, q_len, kv_len;
  int i0;        // first query token of the pair (position within the sequence)
  int j_begin;   // first 64-key step of this split
  int nt[2];     // steps walked for tile t (local step j is global step j_begin + j; interleaved items: j_begin + 2 j + t)
  int n_all;     // ring steps of the item: max(nt[0], nt[1]), or nt[0] + nt[1] for an interleaved item; 0 = nothing to do
  bool il;       // interleaved decode item: one query token, BOTH tiles hold its rows, tile t owns the steps of parity t and the two
                 // (m, l, O) states are merged in the epilogue - a decode row keeps both halves of the CTA busy instead of one
  bool direct[2];  // tile t sees all its keys in this one split: its rows go straight to `out`, not to the fp32 partials
  int blk0, n_pages;
};

// Decoding an item is two levels of dependent global loads (plan entry -> sequence metadata): p2_item_head issues the first,
// p2_item_body the second and the arithmetic, so a role can put other work between them (the softmax warps overlap them with
// the wait for the last P.V and with their epilogue).
__device__ __forceinline__ int p2_item_head(const P2Args& a, int k, P2Item& it) {
  int pair;
  if (a.work_items != nullptr) {
    it.sp = k % a.n_splits;
    it.kvh = (k / a.n_splits) % a.n_kv_heads;
    const int item = k / (a.n_splits * a.n_kv_heads);
    it.b = __ldg(a.work_items + 2 * item);
    pair = __ldg(a.work_items + 2 * item + 1);
  } else {
    it.sp = k % a.n_splits;
    const int slot = (k / a.n_splits) % a.max_pairs;
    it.kvh = (k / (a.n_splits * a.max_pairs)) % a.n_kv_heads;
    it.b = k / (a.n_splits * a.max_pairs * a.n_kv_heads);
    // Sequence-major on purpose: CTAs that run side by side work on the same sequence and share its K/V in L2.  (Slot-major -
    // every sequence's heaviest pair first - was tried: plan-less prefill -6 %, but the vision towers +8 % and the mixed batch
    // +13 %.)
    pair = -1 - slot;  // resolved in the body: the sequence's latest (heaviest) pair first
  }
  return pair;
}

// The sequence metadata of an item as loaded; p2_item_loads issues the loads, p2_item_finish does the arithmetic, so that a role can
// put its own work between the two (the softmax warps: their whole epilogue - the loads used to be consumed right behind the
// wait for the last P.V, a global-load latency in front of the O read-out of every item).
struct P2Raw {
  int q0, q1, k0, k1, c0, c1;
};

template <bool VL = false>
__device__ __forceinline__ void p2_item_loads(const P2Args& a, const P2Item& it, P2Raw& w) {
  w.q0 = __ldg(a.q_cu + it.b);
  w.q1 = __ldg(a.q_cu + it.b + 1);
  w.k0 = __ldg(a.kv_cu + it.b);
  w.k1 = __ldg(a.kv_cu + it.b + 1);
  if constexpr (VL) {
    w.c0 = w.c1 = 0;
  } else {
    w.c0 = __ldg(a.cu_blocks + it.b);
    w.c1 = __ldg(a.cu_blocks + it.b + 1);
  }
}

template <bool VL = false>
__device__ __forceinline__ void p2_item_finish(const P2Args& a, int pair, const P2Raw& w, P2Item& it) {
  it.q_start = w.q0;
  it.q_len = w.q1 - w.q0;
  it.kv_len = w.k1 - w.k0;
  if constexpr (VL) {
    it.blk0 = w.k0;  // first K/V row of the sequence
    it.n_pages = (it.kv_len + a.block_size - 1) / a.block_size;
  } else {
    it.blk0 = w.c0;
    it.n_pages = w.c1 - w.c0;
  }
  const int pair_tokens = 2 * a.tq;
  if (pair < 0) pair += (it.q_len + pair_tokens - 1) / pair_tokens;  // n_pairs - 1 - slot
  it.i0 = pair * pair_tokens;
  it.j_begin = it.sp * a.tiles_per_split;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int first = it.i0 + t * a.tq;
    it.direct[t] = true;
    if (pair < 0 || first >= it.q_len) {
      it.nt[t] = 0;
    } else {
      const int i_last = min(it.q_len, first + a.tq) - 1;
      const int kv_end = (VL && !a.causal) ? it.kv_len : i_last + (it.kv_len - it.q_len) + 1;  // keys [0, kv_end) are visible to the tile
      const int n_vis = (kv_end + kP2TileN - 1) / kP2TileN;
      it.direct[t] = n_vis <= a.direct_limit;  // the same rule merge_partials_kernel applies (direct_tile_tokens, direct_tiles)
      it.nt[t] = it.direct[t] ? (it.sp == 0 ? n_vis : 0) : max(0, min(n_vis - it.j_begin, a.tiles_per_split));
    }
  }
  it.n_all = max(it.nt[0], it.nt[1]);
  it.il = false;
  if constexpr (!VL) {
    if (a.interleave_decode && it.q_len == 1 && it.nt[0] >= 2) {
      it.il = true;
      it.n_all = it.nt[0];
      it.nt[1] = it.nt[0] >> 1;          // odd steps
      it.nt[0] = it.nt[0] - it.nt[1];    // even steps
      it.direct[1] = it.direct[0];
    }
  }
}

template <bool VL = false>
__device__ __forceinline__ void p2_item_body(const P2Args& a, int pair, P2Item& it) {
  P2Raw w;
  p2_item_loads<VL>(a, it, w);
  p2_item_finish<VL>(a, pair, w, it);
}

template <bool VL = false>
__device__ __forceinline__ void p2_decode_item(const P2Args& a, int k, P2Item& it) {
  const int pair = p2_item_head(a, k, it);
  p2_item_body<VL>(a, pair, it);
}

// Work distribution.  Items are claimed DYNAMICALLY from a global counter (the list is sorted heaviest first by the host
// plan, so this is longest-processing-time-first scheduling); warp 8 claims them just in time (kP2ClaimAhead steps before
// it has loaded the current item) and publishes the indices to the other roles through a 4-entry shared-memory ring so that every role
// walks the same sequence while running up to a few items apart.  Without a counter (no workspace) the assignment is
// static in boustrophedon order: even rounds left to right, odd rounds right to left, which pairs a heavy item of one
// round with a light one of the next.  An index >= n_items ends the walk.
__device__ __forceinline__ int p2_claim_item(const P2Args& a, int round, int lane) {
  if (a.work_counter != nullptr) {
    int k = 0;
    if (lane == 0) k = static_cast<int>(atomicAdd(a.work_counter, 1u));
    return __shfl_sync(0xffffffffu, k, 0);
  }
  const int g = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  if (round * g >= a.n_items) return a.n_items;
  const int k = round * g + ((round & 1) ? g - 1 - c : c);
  return k < a.n_items ? k : a.n_items + 1;  // + 1: a hole in the last round, not the end (static mode only)
}

// Timeline instrumentation (dev builds with -DHI_PAIR_TRACE, tools/pair_trace.py): CTA 0 records, per role, clock64 stamps
// of its hand-off points.  Record = clock (40 bits) | tag << 40 | step << 44 | item << 56.
#ifdef HI_PAIR_TRACE
constexpr int kTraceRoles = 12, kTraceCap = 4096;  // K TMA, V TMA, MMA 0 / 1, lane 0 of the 8 softmax warps
static __device__ unsigned long long g_pair_trace[kTraceRoles * kTraceCap];
static __device__ unsigned int g_pair_trace_n[kTraceRoles];
// Per CTA (MMA warp of tile 0): {globaltimer at entry, globaltimer after the last item, items walked, 64-key steps walked}.
static __device__ unsigned long long g_pair_cta_stats[1024 * 4];
__device__ __forceinline__ unsigned long long trace_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#endif

// PF = how many of every 4 (pairs of) exponentials run on the FMA pipes (exp2_poly2) instead of MUFU.EX2.
// MODE 0: paged K/V, head_dim 128 (everything about the head dim is a compile-time constant).
// MODE 1: un-paged varlen K/V (hi_varlen_attention), head_dim any multiple of 8 up to 128 (VL).
// MODE 2: paged K/V, head_dim a multiple of 16 below 128 (64, 96: the reference's FA2 covers 64 / 96 / 128 / 256,
//         csrc/kernel/flash_attn/src/static_switch.h:70-85): the tensor maps carry the real head_dim, TMA zero-fills the dims
//         beyond it, Q.K^T walks head_dim / 16 k-steps and P.V produces head_dim columns, exactly as MODE 1 does.
template <typename T, int NK, int NV, int PF, int MODE = 0>
__global__ void __launch_bounds__(kP2Threads, 1)
paged_attn_pair_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                       const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const P2Args a) {
  using L = P2Smem<NK, NV>;
  static_assert(NK >= 3 && NV >= 3 && NK + NV == 8, "the K and V rings share 128 KiB; each needs 2 steps of lookahead at least");
  constexpr bool VL = MODE == 1;   // un-paged keys, optional causal mask
  constexpr bool VD = MODE != 0;   // head_dim is a run-time value
  // ring stage / phase of running step g (division by a compile-time constant)
  auto k_stage = [](uint32_t g) -> uint32_t { return g % NK; };
  auto k_phase = [](uint32_t g) -> uint32_t { return (g / NK) & 1u; };
  auto v_stage = [](uint32_t g) -> uint32_t { return g % NV; };
  auto v_phase = [](uint32_t g) -> uint32_t { return (g / NV) & 1u; };
  constexpr bool kBf16 = !std::is_same<T, __half>::value;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bars = smem_base + L::kBars;
  auto bar = [&](int idx) -> uint32_t { return bars + idx * 8; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kTmemPtr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pages_per_tile = kP2TileN / a.block_size;
#ifdef HI_PAIR_TRACE
  const int tr_role = (warp == 8) ? 0 : (warp == 11) ? 1 : (warp == 9) ? 2 : (warp == 10) ? 3 : 4 + warp;
  const bool tr_on = blockIdx.x == 0 && lane == 0;
  unsigned int tr_n = 0;
  auto trace = [&](int tag, unsigned int item, int step) {
    if (tr_on && tr_n < kTraceCap)
      g_pair_trace[tr_role * kTraceCap + tr_n++] = (static_cast<unsigned long long>(clock64()) & 0xffffffffffull) | (static_cast<unsigned long long>(tag & 15) << 40) |
                                                   (static_cast<unsigned long long>(step & 4095) << 44) | (static_cast<unsigned long long>(item & 255u) << 56);
  };
#else
  auto trace = [&](int, unsigned int, int) {};
#endif

  // ---- one-time setup ------------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < L::kNumBars; ++i) ptx::mbar_init(bar(i), 1);
    for (int i = 0; i < NK; ++i) ptx::mbar_init(bar(L::bKEmpty + i), 2);  // released by both MMA warps
    for (int i = 0; i < NV; ++i) ptx::mbar_init(bar(L::bVEmpty + i), 2);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(bar(L::bPFull + i), kP2TileM);  // every softmax thread of the tile arrives
    for (int i = 0; i < 2; ++i) ptx::mbar_init(bar(L::bOEmpty + i), kP2TileM);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(bar(L::bItemEmpty + i), 3 + 2 * kP2TileM);  // warps 9, 10, 11 + 256 softmax threads
    ptx::fence_mbar_init();
  }
  if (warp == 8) {
    if (lane == 0) {
      ptx::prefetch_tensormap(&tm_q);
      ptx::prefetch_tensormap(&tm_k);
      ptx::prefetch_tensormap(&tm_v);
      ptx::prefetch_tensormap(&tm_o);
    }
    __syncwarp();
    ptx::tmem_alloc(smem_base + L::kTmemPtr, kP2TmemCols);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // The setup above ran beside the tail of the kernel in front (append / device plan); nothing a kernel produced - q, the
  // appended K / V rows, the device-built plan, the work counter - has been touched yet.
  pdl_wait();
  pdl_launch_dependents();
  volatile int* item_ring = reinterpret_cast<volatile int*>(smem_gen + L::kItemRing);
  // Next work item of a consumer role (see p2_claim_item): entry n of the ring sequence.  `warp_role` consumers stay
  // converged and release the slot from one lane; softmax threads release it individually.
  auto next_item = [&](uint32_t n, bool warp_role) -> int {
    const uint32_t slot = n & 3u;
    ptx::mbar_wait(bar(L::bItemFull + slot), (n >> 2) & 1u);
    const int k = item_ring[slot];
    if (warp_role) {
      __syncwarp();
      if (ptx::elect_one()) ptx::mbar_arrive(bar(L::bItemEmpty + slot));
      __syncwarp();
    } else {
      ptx::mbar_arrive(bar(L::bItemEmpty + slot));
    }
    return k;
  };

  // Running counters (identical in every role, never reset): g = K/V steps so far (ring stage g % N, phase (g / N) & 1);
  // per tile, steps so far (S/P buffer & 1, phase >> 1), items in which the tile was active (Q/O barriers) and items
  // with at least two steps (P.V(n-2) barrier).
  if (warp >= 8) {
    // ========================== TMA producers (warps 8, 11) and MMA warps (9, 10) ====================================
    ptx::setmaxnreg_dec<88>();  // 4 x 88 + 8 x 208 registers per lane = 2016 of 2048; at 72 the MMA-issue loop spills once it tracks interleaved items
    if (warp == 8 || warp == 11) {
      // ---- TMA producer: warp 8 stages Q and the K ring, warp 11 the V ring.  The warp stays converged; one elected lane
      // issues.
      const bool is_k = warp == 8;
      const CUtensorMap* tm = is_k ? &tm_k : &tm_v;
      const uint32_t page_half_bytes = static_cast<uint32_t>(a.block_size) * 128u;
      const uint32_t n_halves = VD ? static_cast<uint32_t>(a.n_halves) : 2u;
      const int b_full = is_k ? L::bKFull : L::bVFull;
      const int b_empty = is_k ? L::bKEmpty : L::bVEmpty;
      const uint32_t ring = smem_base + (is_k ? L::kK : L::kV);
      uint32_t g = 0, n_q[2] = {0, 0}, n_tail = 0, n_it = 0;
      int k_next = is_k ? p2_claim_item(a, 0, lane) : 0;
      for (;;) {
        int k;
        if (is_k) {  // publish the current item to the other roles (the next one is claimed near the end of this one's K loop)
          k = k_next;
          const uint32_t slot = n_it & 3u;
          ptx::mbar_wait(bar(L::bItemEmpty + slot), ((n_it >> 2) & 1u) ^ 1u);
          if (ptx::elect_one()) {
            item_ring[slot] = k;
            ptx::mbar_arrive(bar(L::bItemFull + slot));
          }
          __syncwarp();
        } else {
          k = next_item(n_it, true);
        }
        ++n_it;
        if (k >= a.n_items) break;
        P2Item it;
        p2_decode_item<VL>(a, k, it);
        if (it.n_all == 0) {
          if (is_k) k_next = p2_claim_item(a, static_cast<int>(n_it), lane);
          continue;
        }
        trace(1, n_it, it.n_all);
        if (is_k) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (it.nt[t] > 0) {
              ptx::mbar_wait(bar(L::bQEmpty + t), (n_q[t] & 1u) ^ 1u);  // the previous item's Q_t.K products are done
              trace(2, n_it, t);
              if (ptx::elect_one()) {
                const uint32_t dst = smem_base + L::kQ + t * kP2QTile;
                ptx::mbar_arrive_expect_tx(bar(L::bQFull + t), n_halves * static_cast<uint32_t>(a.group * a.tq) * 128u);
                const int q_tok = it.q_start + it.i0 + (it.il ? 0 : t * a.tq);  // interleaved items: both tiles hold the same rows
                ptx::tma_load_3d(dst, &tm_q, bar(L::bQFull + t), 0, it.kvh * a.group, q_tok);
                if (n_halves == 2u) ptx::tma_load_3d(dst + kP2QHalf, &tm_q, bar(L::bQFull + t), 64, it.kvh * a.group, q_tok);
              }
              __syncwarp();
              ++n_q[t];
            }
          }
        }
        // Lane p holds the block id of page p of the step; the ids of step j+1 are fetched while step j is being issued.
        auto pages_of = [&](int j, int& n_valid) -> int {
          const int page0 = (it.j_begin + j) * pages_per_tile;
          n_valid = (a.debug & 4) ? 0 : max(0, min(pages_per_tile, it.n_pages - page0));
          if constexpr (VL) return 0;
          return (j < it.n_all && lane < n_valid) ? __ldg(a.block_tables + it.blk0 + page0 + lane) : 0;
        };
        int n_valid_next = 0;
        int blk_next = pages_of(0, n_valid_next);
        for (int j = 0; j < it.n_all; ++j, ++g) {
          // The next item is claimed kP2ClaimAhead steps before this warp is done with the current one - early enough to hide
          // the atomic and the decode, late enough that the claim order follows who really finishes first.  (Claiming at the
          // start of an item makes the first two rounds of a launch static: every CTA takes two of the heaviest items at t = 0,
          // measured on the config-3 mixed batch as CTAs ending between 126 and 208 us.)
          if (is_k && j == max(it.n_all - kP2ClaimAhead, 0)) k_next = p2_claim_item(a, static_cast<int>(n_it), lane);
          const int n_valid = n_valid_next;
          const int blk_lane = blk_next;
          blk_next = pages_of(j + 1, n_valid_next);
          const uint32_t tx = static_cast<uint32_t>(n_valid) * n_halves * page_half_bytes;
          const int st = static_cast<int>(is_k ? k_stage(g) : v_stage(g));
          const uint32_t ph = is_k ? k_phase(g) : v_phase(g);
          const int kv0 = (it.j_begin + j) * kP2TileN;
          const bool tail = !is_k && (kv0 + kP2TileN > it.kv_len);  // at most one such step per item
          ptx::mbar_wait(bar(b_empty + st), ph ^ 1u);
          trace(3, n_it, j);
          const uint32_t full_bar = tail ? bar(L::bVTail) : bar(b_full + st);
          uint32_t dst = ring + st * kP2Tile;
          if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(full_bar, tx);
          __syncwarp();
          for (int p = 0; p < n_valid; ++p, dst += page_half_bytes) {
            const int slot0 = VL ? it.blk0 + ((it.j_begin + j) * pages_per_tile + p) * a.block_size
                                 : __shfl_sync(0xffffffffu, blk_lane, p) * a.block_size;
            if (ptx::elect_one()) {
              ptx::tma_load_3d(dst, tm, full_bar, 0, it.kvh, slot0);
              if (n_halves == 2u) ptx::tma_load_3d(dst + kP2Half, tm, full_bar, 64, it.kvh, slot0);
            }
            __syncwarp();
          }
          if (tail) {
            // Keys at or beyond kv_len carry P == 0, but their V rows are whatever the pool / stale shared memory holds:
            // zero them (0 * NaN must not reach O), then publish the step.
            ptx::mbar_wait(bar(L::bVTail), n_tail & 1u);
            ++n_tail;
            uint8_t* vt = smem_gen + L::kV + st * kP2Tile;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int r = max(0, it.kv_len - kv0) + lane; r < kP2TileN; r += 32) {
              uint4* row0 = reinterpret_cast<uint4*>(vt + r * 128);
              uint4* row1 = reinterpret_cast<uint4*>(vt + kP2Half + r * 128);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                row0[e] = z;
                row1[e] = z;
              }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (ptx::elect_one()) ptx::mbar_arrive(bar(L::bVFull + st));
            __syncwarp();
          }
        }
      }
    } else {
      // ---- MMA warp of tile t: the warp stays converged (every lane polls the barriers), one elected lane issues ------------
      const int t = warp - 9;
      constexpr uint32_t idesc_qk = ptx::make_idesc_f16(kBf16, false, false, kP2TileM, kP2TileN);
      const uint32_t idesc_pv = VD ? a.idesc_pv : ptx::make_idesc_f16(kBf16, false, true, kP2TileM, kP2D);
      const int n_kk = VD ? a.n_kk : 8;
      // Descriptors of the operand bases, built once; stages and k-steps only add to the 14-bit start-address field.
      const uint64_t desc_q = ptx::make_smem_desc_sw128(smem_base + L::kQ + t * kP2QTile, 16, 1024);
      const uint64_t desc_k = ptx::make_smem_desc_sw128(smem_base + L::kK, 16, 1024);
      const uint64_t desc_v = ptx::make_smem_desc_sw128(smem_base + L::kV, kP2Half, 1024);
      const uint32_t tmem_t = tmem_base + t * kP2ColTile;
      const uint32_t tmem_o = tmem_t + kP2ColO;
      const bool no_mma = (a.debug & 2) != 0;
      auto issue_qk = [&](int st, int buf, bool last_of_item) {  // S_t[buf] = Q_t . K(stage)^T; releases the K stage
        const uint64_t dk = desc_k + static_cast<uint64_t>(st * (kP2Tile >> 4));
        const uint32_t tmem_s = tmem_t + buf * kP2TileN;
        if (!no_mma) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 2 halves x 4 k-steps of 16 dims; K-major operands, 8-row groups 1024 B apart
            if (VD && kk >= n_kk) break;
            ptx::mma_f16_ss(tmem_s, desc_q + static_cast<uint64_t>(((kk >> 2) * kP2QHalf + (kk & 3) * 32) >> 4),
                            dk + static_cast<uint64_t>(((kk >> 2) * kP2Half + (kk & 3) * 32) >> 4), idesc_qk, kk > 0);
          }
        }
        ptx::mma_commit(bar(L::bSFull + 2 * t + buf));
        ptx::mma_commit(bar(L::bKEmpty + st));
        if (last_of_item) ptx::mma_commit(bar(L::bQEmpty + t));
      };
      auto issue_pv = [&](int st, int buf, bool accumulate) {  // O_t (+)= P_t[buf] . V(stage); releases the V stage
        const uint64_t dv = desc_v + static_cast<uint64_t>(st * (kP2Tile >> 4));
        const uint32_t tmem_p = tmem_t + buf * kP2TileN;
        if (!no_mma) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 k-steps of 16 keys; A = P in TMEM (8 columns per step), B = V MN-major
            ptx::mma_f16_ts(tmem_o, tmem_p + kk * 8, dv + static_cast<uint64_t>((kk * 2048) >> 4), idesc_pv, accumulate || (kk > 0));
          }
        }
        ptx::mma_commit(bar(L::bVEmpty + st));
      };
      uint32_t g = 0, gs = 0, n_act = 0, n_it = 0;
#ifdef HI_PAIR_TRACE
      const unsigned long long cta_t0 = trace_globaltimer();
      unsigned int cta_items = 0;
#endif
      // The next item is decoded during the last steps of the current one: plan entry at the top of step n-3, sequence metadata
      // loads at the top of step n-2, the arithmetic behind the last P.V - so no load latency sits between a P and its P.V.
      P2Item nxt;
      P2Raw raw_nxt;
      bool more_items = false;
      int pair_nxt = 0;
      {
        const int k = next_item(n_it++, true);
        more_items = k < a.n_items;
        if (more_items) p2_decode_item<VL>(a, k, nxt);
      }
      while (more_items) {
        const P2Item it = nxt;
        if (it.n_all == 0) {
          const int k = next_item(n_it++, true);
          more_items = k < a.n_items;
          if (more_items) p2_decode_item<VL>(a, k, nxt);
          continue;
        }
        const int n_t = t ? it.nt[1] : it.nt[0];
        const int n_all = it.n_all;
        // Which ring steps this tile consumes, and its running index over them.  Ordinary items: steps 0 .. n_t - 1, Q.K^T issued two
        // steps ahead of P.V.  Interleaved decode items: the steps of parity t (own index j >> 1); the look-ahead of two ring steps is
        // then ONE own step, so S(u + 1) is issued right behind P.V(u) and the two tiles alternate on the tensor pipe.
        const int il_sh = it.il ? 1 : 0;     // own index of ring step j = j >> il_sh
        const int il_par = it.il ? t : 0;    // ... and the tile owns it iff (j & il_sh) == il_par (and the index is below n_t)
        auto owns = [&](int j) -> bool { return (j & il_sh) == il_par && (j >> il_sh) < n_t; };
        auto own_idx = [&](int j) -> int { return j >> il_sh; };
        trace(1, n_it, n_t);
        // prologue: S_t(0) and S_t(1).  Steps at or beyond n_t (this tile sees fewer keys than its sibling) only release
        // the ring stages.
        if (n_t > 0) ptx::mbar_wait(bar(L::bQFull + t), n_act & 1u);
        trace(2, n_it, 0);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          if (jj < n_all) {
            const uint32_t gj = g + jj;
            const int ks = static_cast<int>(k_stage(gj));
            ptx::mbar_wait(bar(L::bKFull + ks), k_phase(gj));
            ptx::tc_fence_after_sync();
            if (ptx::elect_one()) {
              if (owns(jj)) issue_qk(ks, (gs + own_idx(jj)) & 1, own_idx(jj) == n_t - 1); else ptx::mbar_arrive(bar(L::bKEmpty + ks));
            }
            __syncwarp();
          }
        }
        // step j: P_t.V(j), then Q_t.K(j+2) into the S buffer P_t(j) just vacated (in order behind P_t.V(j))
        for (int j = 0; j < n_all; ++j) {
          const uint32_t gj = g + j;
          const int u = own_idx(j);  // own index of step j (meaningful when this tile owns it)
          const int s = static_cast<int>(v_stage(gj)), s2 = static_cast<int>(k_stage(gj + 2)), buf = (gs + u) & 1;
          const bool has_pv = owns(j);
          const bool more = j + 2 < n_all;
          if (j == max(n_all - 3, 0)) {  // warp 8 published the next index when it had issued this item's last K load: the K this step's Q.K(j + 2) needs
            const int k = next_item(n_it++, true);
            more_items = k < a.n_items;
            if (more_items) pair_nxt = p2_item_head(a, k, nxt);
          }
          if (j == max(n_all - 2, 0) && more_items) p2_item_loads<VL>(a, nxt, raw_nxt);  // consumed behind the last P.V (below the loop)
          ptx::mbar_wait(bar(L::bVFull + s), v_phase(gj));
          if (more) ptx::mbar_wait(bar(L::bKFull + s2), k_phase(gj + 2));
          if (has_pv) {
            if (u == 0) ptx::mbar_wait(bar(L::bOEmpty + t), (n_act & 1u) ^ 1u);  // the previous item's O_t has been read out
            ptx::mbar_wait(bar(L::bPFull + 2 * t + buf), ((gs + u) >> 1) & 1u);
          }
          ptx::tc_fence_after_sync();
          trace(3, n_it, j);
          if (ptx::elect_one()) {
            if (has_pv) {
              issue_pv(s, buf, u > 0);
              if (u == n_t - 2) ptx::mma_commit(bar(L::bPvDone + t));
              if (u == n_t - 1) ptx::mma_commit(bar(L::bOFull + t));
            } else {
              ptx::mbar_arrive(bar(L::bVEmpty + s));
            }
            if (more) {
              const int u2 = own_idx(j + 2);
              if (owns(j + 2)) issue_qk(s2, (gs + u2) & 1, u2 == n_t - 1); else ptx::mbar_arrive(bar(L::bKEmpty + s2));
            }
          }
          __syncwarp();
          trace(4, n_it, j);
        }
        // (the arithmetic of the decode - a global-load latency when it sat at the top of the last step, in front of the P.V every
        // softmax thread of the tile is waiting for - runs here, off the chain last P -> P.V -> O read-out)
        if (more_items) p2_item_finish<VL>(a, pair_nxt, raw_nxt, nxt);
        g += n_all;
        gs += n_t;
        if (n_t > 0) ++n_act;
#ifdef HI_PAIR_TRACE
        ++cta_items;
#endif
      }
#ifdef HI_PAIR_TRACE
      if (t == 0 && lane == 0 && blockIdx.x < 1024) {
        g_pair_cta_stats[blockIdx.x * 4 + 0] = cta_t0;
        g_pair_cta_stats[blockIdx.x * 4 + 1] = trace_globaltimer();
        g_pair_cta_stats[blockIdx.x * 4 + 2] = cta_items;
        g_pair_cta_stats[blockIdx.x * 4 + 3] = g;
      }
#endif
    }
  } else {
    // ================================================ softmax + epilogue ===========================================
    ptx::setmaxnreg_inc<208>();
    const int t = warp >> 2;                         // query tile of this warpgroup
    const int r = threadIdx.x & 127;                 // tile row == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_t = tmem_base + lane_base + t * kP2ColTile;
    const uint32_t tmem_o = tmem_t + kP2ColO;
    const int tok = r / a.group;
    const int g = r - tok * a.group;
    uint32_t gs = 0, n_act = 0, n_pv2 = 0, n_it = 0;

    // The NEXT item is decoded while this one finishes: its index is fetched after the last step (warp 8 published it several
    // steps ago), the plan entry is loaded behind the wait for the last P.V and the sequence metadata behind the epilogue, so
    // the two levels of global-load latency are off the path between two items' softmax work.
    P2Item nxt;
    bool more = false;
    {
      const int k = next_item(n_it++, false);
      more = k < a.n_items;
      if (more) p2_decode_item<VL>(a, k, nxt);
    }
    while (more) {
      const P2Item it = nxt;
      const int n_mine = t ? it.nt[1] : it.nt[0];
      if (n_mine == 0) {
        const int k = next_item(n_it++, false);
        more = k < a.n_items;
        if (more) p2_decode_item<VL>(a, k, nxt);
        continue;
      }
      trace(1, n_it, n_mine);
      const bool il = it.il;                           // interleaved decode item: both tiles hold the token, tile t walks the steps of parity t
      const int first = it.i0 + (il ? 0 : t * a.tq);
      const int i = first + tok;                       // query position within the sequence
      const bool row_valid = (tok < a.tq) && (i < it.q_len);
      const int lim = (VL && !a.causal) ? it.kv_len - 1 : i + (it.kv_len - it.q_len);  // last visible key index of this row
      float m_used = 0.f;                              // exponent reference (scaled log2 domain)
      float l = 0.f;
      // Rows past the tile's valid (token, head) pairs are padding; a warp that owns only padding rows skips the math.
      const int rows_real = (min(it.q_len, first + a.tq) - first) * a.group;
      const bool warp_active = (warp & 3) * 32 < rows_real;

      for (int j = 0; j < n_mine; ++j) {
        const uint32_t sj = gs + j;
        const int buf = sj & 1;
        const uint32_t tmem_s = tmem_t + buf * kP2TileN;
        const int kv0 = (it.j_begin + (il ? 2 * j + t : j)) * kP2TileN;
        const int col_lim = lim - kv0;                 // columns [0, col_lim] are visible
        ptx::mbar_wait(bar(L::bSFull + 2 * t + buf), (sj >> 1) & 1u);
        ptx::tc_fence_after_sync();
        trace(3, n_it, j);
        if (warp_active && !(a.debug & 1)) {
          uint32_t s[2][32];
#pragma unroll
          for (int c = 0; c < 2; ++c) ptx::tmem_ld_x32(tmem_s + c * 32, s[c]);
          ptx::tmem_wait_ld();
          if (__any_sync(0xffffffffu, col_lim < kP2TileN - 1)) {  // step touches the causal diagonal / end of the sequence
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (c * 32 + e > col_lim) s[c][e] = 0xff800000u;  // -inf
            }
          }
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                mx4[u] = fmax3(mx4[u], __uint_as_float(s[c][e + 2 * u]), __uint_as_float(s[c][e + 2 * u + 1]));
            }
          }
          const float mxs = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * a.scale_log2;
          if (j == 0) {
            m_used = (mxs == -INFINITY) ? 0.f : mxs;
          } else if (__any_sync(0xffffffffu, r < rows_real && mxs > m_used + kP2Rescale)) {
            // (padding rows hold scores of stale Q rows: they must not vote, or the warp's rounding would depend on which
            // item used this buffer before - results would stay within tolerance but not be reproducible run to run)
            // Lazy rescale: the warp pays the TMEM round trip only when some row's max grew by more than 2^8.  O_t must be
            // quiescent: P_t.V(j-1) complete (implied by S_t(j+1), issued behind it; the last step has its own commit) and
            // P_t.V(j) not issued before this thread's P arrival below.
            // (interleaved items: Q.K(j) was issued right behind P.V(j - 1), so the S(j) this thread already waited for implies it)
            trace(7, n_it, j);
            if (!il) {
              if (j + 1 < n_mine) ptx::mbar_wait(bar(L::bSFull + 2 * t + (buf ^ 1)), ((sj + 1) >> 1) & 1u);
              else ptx::mbar_wait(bar(L::bPvDone + t), n_pv2 & 1u);
            }
            ptx::tc_fence_after_sync();
            const float m_new = fmaxf(m_used, mxs);
            const float alpha = fast_exp2(m_used - m_new);
            l *= alpha;
            m_used = m_new;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              ptx::tmem_ld_x32(tmem_o + c * 32, o);
              ptx::tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              ptx::tmem_st_x32(tmem_o + c * 32, o);
            }
          }
          // P = exp2(S * scale - m) as 16-bit pairs, written over S (every S column is already in registers)
          const float2 sc2 = make_float2(a.scale_log2, a.scale_log2);
          const float2 nm2 = make_float2(-m_used, -m_used);
          float2 ls2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              const float2 x2 = ffma2(make_float2(__uint_as_float(s[c][e]), __uint_as_float(s[c][e + 1])), sc2, nm2);
              const float2 p2 = (((e >> 1) & 3) >= 4 - PF) ? exp2_poly2(x2) : make_float2(fast_exp2(x2.x), fast_exp2(x2.y));
              ls2[(e >> 1) & 1] = fadd2(ls2[(e >> 1) & 1], p2);
              pk[e >> 1] = pack2<T>(p2.x, p2.y);
            }
            ptx::tmem_st_x16(tmem_s + c * 16, pk);
          }
          l += (ls2[0].x + ls2[0].y) + (ls2[1].x + ls2[1].y);
        }
        ptx::tmem_wait_st();
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(bar(L::bPFull + 2 * t + buf));
        trace(4, n_it, j);
      }

      // ---- epilogue: O / l -> out (or the fp32 split-KV partial) ---------------------------------------------------------
      int pair_nxt = 0;
      {
        const int k = next_item(n_it++, false);
        more = k < a.n_items;
        if (more) pair_nxt = p2_item_head(a, k, nxt);
      }
      P2Raw raw_nxt;
      if (more) p2_item_loads<VL>(a, nxt, raw_nxt);  // needs the plan entry: that latency and these loads' run behind the wait for the last P.V
      ptx::mbar_wait(bar(L::bOFull + t), n_act & 1u);
      ptx::tc_fence_after_sync();
      trace(5, n_it, 0);
      // Interleaved decode item: tile 1 hands its (m, l, O) to tile 0 through shared memory (tile 1's staging buffer, rows of
      // kP2Xchg floats: O, m, l) and is done; tile 0 folds that state into its own O in TMEM and then leaves through the ordinary epilogue
      // below.  Two CTA-wide rendezvous per item: "tile 1's state is in shared memory" and "tile 0 has read it" (the buffer is
      // tile 1's TMA staging buffer for its next whole tile).  A separate pass on purpose: the epilogue below keeps its
      // register footprint (see the note on spills in DESIGN.md).
      if (il) {
        float* xchg = reinterpret_cast<float*>(smem_gen + L::kOst + kP2QHalf) + r * kP2Xchg;  // 16-byte aligned rows
        if (t == 1) {
          if (r == 0) ptx::bulk_wait_group_read<0>();  // this tile's last TMA store has read the buffer
          ptx::named_bar_sync(2, kP2TileM);
          if (warp_active) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t v[32];
              ptx::tmem_ld_x32(tmem_o + c * 32, v);
              ptx::tmem_wait_ld();
              if (row_valid) {
                float4* x4 = reinterpret_cast<float4*>(xchg + c * 32);
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                  x4[e >> 2] = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
              }
            }
            if (row_valid) {
              xchg[kP2D] = m_used;
              xchg[kP2D + 1] = l;
            }
          }
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive(bar(L::bOEmpty + t));       // O_1 has left TMEM
          ptx::named_bar_sync(3, 2 * kP2TileM);        // tile 1's state is in shared memory
          ptx::named_bar_sync(3, 2 * kP2TileM);        // tile 0 has read it
          trace(6, n_it, 0);
          if (more) p2_item_finish<VL>(a, pair_nxt, raw_nxt, nxt);
          gs += n_mine;
          ++n_act;
          if (n_mine >= 2) ++n_pv2;
          continue;
        }
        ptx::named_bar_sync(3, 2 * kP2TileM);          // tile 1's state is in shared memory
        if (warp_active) {
          float w_own = 1.f, w_peer = 0.f;
          if (row_valid) {
            const float m_peer = xchg[kP2D], l_peer = xchg[kP2D + 1];
            const float m_new = fmaxf(m_used, m_peer);
            w_own = fast_exp2(m_used - m_new);
            w_peer = fast_exp2(m_peer - m_new);
            l = l * w_own + l_peer * w_peer;
            m_used = m_new;
          }
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_x32(tmem_o + c * 32, v);
            ptx::tmem_wait_ld();
            if (row_valid) {
              const float4* x4 = reinterpret_cast<const float4*>(xchg + c * 32);
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                const float4 p = x4[e >> 2];
                v[e] = __float_as_uint(__uint_as_float(v[e]) * w_own + p.x * w_peer);
                v[e + 1] = __float_as_uint(__uint_as_float(v[e + 1]) * w_own + p.y * w_peer);
                v[e + 2] = __float_as_uint(__uint_as_float(v[e + 2]) * w_own + p.z * w_peer);
                v[e + 3] = __float_as_uint(__uint_as_float(v[e + 3]) * w_own + p.w * w_peer);
              }
            }
            ptx::tmem_st_x32(tmem_o + c * 32, v);
          }
          ptx::tmem_wait_st();
        }
        ptx::named_bar_sync(3, 2 * kP2TileM);          // tile 0 has read the exchange buffer
      }
      const float inv_l = 1.f / l;
      const int head = it.kvh * a.group + g;
      const int d_out = VD ? a.head_dim : kP2D;
      T* orow = static_cast<T*>(a.out) + static_cast<int64_t>(it.q_start + i) * a.out_row_stride + head * d_out;
      const int64_t pidx = (static_cast<int64_t>(it.q_start + i) * a.n_qo_heads + head) * a.n_splits + it.sp;
      const bool direct = t ? it.direct[1] : it.direct[0];
      if (!direct && row_valid) {
        a.part_ml[pidx * 2 + 0] = m_used;
        a.part_ml[pidx * 2 + 1] = l;
      }
      // Whole tiles of a head_dim-128 launch whose rows go straight to `out` leave through shared memory and ONE TMA store
      // per 64-dim half: the thread (= row) writes its 128 B of the half into the tile's staging buffer in the 128B-swizzled
      // layout of the tensor map (the layout Q was loaded in) and a bulk tensor store moves the {64 dims, group heads, tq
      // tokens} box.  (A per-row 16-byte store instruction touches 32 different 256-byte rows per warp.)  Tiles cut by the end
      // of the sequence (their box would cover the next sequence's tokens) and other head dims store per row; split-KV
      // partials leave as fp32.  All three share the TMEM loads and the packing: one register footprint for the epilogue.
      const bool tma_out = a.use_tma_store && direct && d_out == kP2D && first + a.tq <= it.q_len;  // uniform over the warpgroup
      const uint32_t stage = smem_base + L::kOst + t * kP2QHalf;
      const uint32_t row_addr = stage + static_cast<uint32_t>(r) * 128u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[32];  // 64 dims of the row as 16-bit pairs
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          if (warp_active && (h * 2 + c) * 32 < d_out) {
            ptx::tmem_ld_x32(tmem_o + (h * 2 + c) * 32, v);
            ptx::tmem_wait_ld();
          }
          if (!direct && warp_active && row_valid) {
            float4* dst = reinterpret_cast<float4*>(a.part_o + pidx * kP2D + (h * 2 + c) * 32);
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              dst[e >> 2] = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
          }
#pragma unroll
          for (int e = 0; e < 32; e += 2) pk[c * 16 + (e >> 1)] = pack2<T>(__uint_as_float(v[e]) * inv_l, __uint_as_float(v[e + 1]) * inv_l);
        }
        if (h == 1) {  // O_t has left TMEM: the next item's first P.V may overwrite it
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive(bar(L::bOEmpty + t));
        }
        if (tma_out) {
          // (the TMEM loads and the packing of half 1 ran while the TMA engine was still reading half 0 out of the buffer)
          if (r == 0) ptx::bulk_wait_group_read<0>();  // the store that last used this buffer has read it
          ptx::named_bar_sync(1 + t, kP2TileM);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            ptx::st_shared_v4(row_addr + static_cast<uint32_t>((c ^ (r & 7)) << 4), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(1 + t, kP2TileM);
          if (r == 0) {
            ptx::tma_store_3d(&tm_o, stage, h * 64, it.kvh * a.group, it.q_start + first);
            ptx::bulk_commit_group();
          }
        } else if (direct && warp_active && row_valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (VD && h * 64 + c * 8 >= d_out) break;
            *reinterpret_cast<uint4*>(orow + h * 64 + c * 8) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
          }
        }
      }
      trace(6, n_it, 0);
      if (more) p2_item_finish<VL>(a, pair_nxt, raw_nxt, nxt);  // the loads were issued before the epilogue
      gs += n_mine;
      ++n_act;
      if (n_mine >= 2) ++n_pv2;
    }
    if (r == 0) ptx::bulk_wait_group<0>();  // the staging buffers must outlive the last TMA store's reads; writes done before exit
  }

  // ---- teardown ----------------------------------------------------------------------------------------------------
#ifdef HI_PAIR_TRACE
  if (tr_on) g_pair_trace_n[tr_role] = tr_n;
#endif
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kP2TmemCols);
  }
}

// ---- plan on the device --------------------------------------------------------------------------------------------------
// Callers at the bare mha_varlen_fwd seam bring no host plan and their sequence lengths live in device memory.  One CTA builds
// the list the host plan would have held: every pair of tiles of every sequence as (sequence, pair), heaviest first
// (counting sort on the number of keys the pair's last token sees, kPlanBuckets buckets, order inside a bucket arbitrary -
// results never depend on the order), padded up to the host-known bound n_slots = n_seqs + n_tokens / pair_tokens with
// entries whose pair index lies beyond any sequence (the kernel decodes those as empty items).  Also zeroes the work counter.
constexpr int kPlanThreads = 1024, kPlanBuckets = 1024;  // 32 buckets per lane of the scanning warp
__global__ void __launch_bounds__(kPlanThreads) p2_plan_kernel(const int32_t* __restrict__ q_cu, const int32_t* __restrict__ kv_cu, int n_seqs,
                                                              int pair_tokens, int max_kv_len, int n_slots, int32_t* __restrict__ work_items,
                                                              unsigned int* __restrict__ work_counter) {
  __shared__ int hist[kPlanBuckets];
  __shared__ int total;
  pdl_wait();  // (a previous attention call on this stream may still be walking the list this kernel rewrites)
  pdl_launch_dependents();
  const int width = max_kv_len / kPlanBuckets + 1;  // keys per bucket
  for (int i = threadIdx.x; i < kPlanBuckets; i += kPlanThreads) hist[i] = 0;
  if (threadIdx.x == 0) {
    total = 0;
    if (work_counter != nullptr) *work_counter = 0u;
  }
  __syncthreads();
  auto bucket_of = [&](int cost) { return min(max(cost, 0) / width, kPlanBuckets - 1); };
  // a warp per sequence, its lanes over the sequence's pairs: one 8k-token prefill and a thousand decode rows both take a
  // few dozen iterations
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = kPlanThreads / 32;
  for (int b = warp; b < n_seqs; b += n_warps) {
    const int q = q_cu[b + 1] - q_cu[b], kv = kv_cu[b + 1] - kv_cu[b];
    const int n_pairs = (q + pair_tokens - 1) / pair_tokens;
    for (int p = lane; p < n_pairs; p += 32) atomicAdd(&hist[bucket_of(kv - q + min(q, (p + 1) * pair_tokens))], 1);
    if (lane == 0) atomicAdd(&total, n_pairs);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan from the heaviest bucket down: hist[k] becomes the first slot of bucket k
    constexpr int kPer = kPlanBuckets / 32;
    const int top = kPlanBuckets - 1 - lane * kPer;  // lane 0 owns the heaviest kPer buckets
    int mine = 0;
#pragma unroll 8
    for (int i = 0; i < kPer; ++i) mine += hist[top - i];
    int before = mine;  // inclusive scan over the lanes, then shifted
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, before, off);
      if (lane >= off) before += up;
    }
    int run = before - mine;
#pragma unroll 8
    for (int i = 0; i < kPer; ++i) {
      const int c = hist[top - i];
      hist[top - i] = run;
      run += c;
    }
  }
  __syncthreads();
  for (int b = warp; b < n_seqs; b += n_warps) {
    const int q = q_cu[b + 1] - q_cu[b], kv = kv_cu[b + 1] - kv_cu[b];
    const int n_pairs = (q + pair_tokens - 1) / pair_tokens;
    for (int p = lane; p < n_pairs; p += 32) {
      const int pos = atomicAdd(&hist[bucket_of(kv - q + min(q, (p + 1) * pair_tokens))], 1);
      if (pos < n_slots) {
        work_items[2 * pos] = b;
        work_items[2 * pos + 1] = p;
      }
    }
  }
  __syncthreads();
  for (int i = total + threadIdx.x; i < n_slots; i += kPlanThreads) {
    work_items[2 * i] = 0;
    work_items[2 * i + 1] = 1 << 20;  // beyond the last pair of any sequence: an empty item
  }
}

bool attn_pair_supported(const HiAttnArgs& args) {
  const int group = args.n_kv_heads > 0 ? args.n_qo_heads / args.n_kv_heads : 0;
  const bool dim_ok = args.head_dim == kP2D || (args.head_dim >= 16 && args.head_dim < kP2D && (args.head_dim % 16) == 0);
  return (args.dtype == HI_F16 || args.dtype == HI_BF16) && dim_ok && group >= 1 && group <= kP2TileM &&
         args.block_size >= 8 && args.block_size <= kP2TileN && (kP2TileN % args.block_size) == 0 &&
         (args.q_row_stride % 8) == 0 && (args.out_row_stride % 8) == 0 && aligned_to(args.q, 16) &&
         aligned_to(args.out, 16) && aligned_to(args.key_cache, 16) && aligned_to(args.value_cache, 16) &&
         args.n_blocks > 0;
}

template <typename T, int PF, int MODE, int NK, int NV>
static int launch_pair_ring(int device, const P2Args& a, const CUtensorMap& mq, const CUtensorMap& mk,
                            const CUtensorMap& mv, const CUtensorMap& mo, cudaStream_t stream) {
  // NK + NV = 8 steps of 64 keys (16 KiB each) + 64 KiB of Q + 32 KiB of output staging = 224 KiB
  using L = P2Smem<NK, NV>;
  static PerDeviceFlags configured;
  HI_CUDA(configure_dynamic_smem(configured, paged_attn_pair_kernel<T, NK, NV, PF, MODE>, L::kDynamicBytes));
  // persistent: one CTA per SM walks the items with a stride of the grid size
  const int n_sms = sm_count_of(device);
  HI_CHECK_ARG(n_sms > 0, "paged_attention: cannot read the SM count of device %d", device);
  int ctas = a.n_items < n_sms ? a.n_items : n_sms;
  if (const char* env = tuning_env("HI_PAIR_CTAS")) ctas = atoi(env) > 0 ? atoi(env) : ctas;  // tuning / test override
  const dim3 grid(ctas, 1, 1);
  timing_mark_start(stream);
  HI_CUDA(launch_pdl(paged_attn_pair_kernel<T, NK, NV, PF, MODE>, grid, dim3(kP2Threads), L::kDynamicBytes, stream, mq, mk, mv, mo, a));
  timing_mark_stop(stream);
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

// Split of the 8 ring stages between K and V: 4 + 4.  (Q.K^T runs two steps ahead of P.V, so 5 + 3 would give both rings the same
// lookahead; measured on B200 it loses 2-6 % on every prefill shape - the V ring then stalls P.V behind the softmax of the
// lagging tile - so only 4 + 4 is instantiated.  The kernel itself is generic in NK / NV.)
template <typename T, int PF, int MODE = 0>
static int launch_pair_t(int device, const P2Args& a, const CUtensorMap& mq, const CUtensorMap& mk,
                         const CUtensorMap& mv, const CUtensorMap& mo, cudaStream_t stream) {
  return launch_pair_ring<T, PF, MODE, 4, 4>(device, a, mq, mk, mv, mo, stream);
}

int launch_attn_pair(const HiAttnArgs& args, cudaStream_t stream) {
  if (!attn_pair_supported(args)) {
    set_error("paged_attention: the tcgen05 pair-tile path needs fp16/bf16, head_dim 128 or a multiple of 16 below it, block_size in {8,16,32,64} and 16-byte aligned rows");
    return HI_ERR_UNSUPPORTED;
  }
  P2Args a{};
  a.out = args.out;
  a.out_row_stride = args.out_row_stride;
  a.q_cu = args.q_cu_seq_lens;
  a.kv_cu = args.kv_cu_seq_lens;
  a.block_tables = args.block_tables;
  a.cu_blocks = args.cu_blocks_lens;
  const bool host_plan = args.work_items != nullptr && args.n_work_items > 0 && args.work_tile_tokens == 2 * (kP2TileM / (args.n_qo_heads / args.n_kv_heads));
  a.work_items = host_plan ? args.work_items : nullptr;
  a.n_qo_heads = args.n_qo_heads;
  a.n_kv_heads = args.n_kv_heads;
  a.group = args.n_qo_heads / args.n_kv_heads;
  a.block_size = args.block_size;
  a.tq = kP2TileM / a.group;
  a.scale_log2 = args.softmax_scale * 1.4426950408889634f;

  // ---- split-KV ---------------------------------------------------------------------------------------------------------
  // One CTA per SM: a launch is balanced when no CTA carries much more than (total work / #SMs).  With the host's work
  // hint the chunk is sized from the real total (ragged batches: a few long sequences among short ones get split, the
  // short ones do not); without it only launches too small to fill the machine are split.
  constexpr int kSms = 148;
  constexpr int kMinTilesPerSplit = 4;   // never finer than 256 tokens
  const int n_pairs = (args.max_q_len + 2 * a.tq - 1) / (2 * a.tq);
  // Without a host plan the list is built on the device (p2_plan_kernel) in the tail of the workspace, in front of the work
  // counter; its length is the host-known bound n_seqs + n_tokens / pair_tokens.
  int64_t n_list = host_plan ? args.n_work_items : 0;
  int32_t* dev_plan = nullptr;
  {
    const char* env = tuning_env("HI_PAIR_DEVICE_PLAN");  // tuning / test override: "0" walks the tiles in sequence order
    const int64_t bound = static_cast<int64_t>(args.n_seqs) + args.n_tokens / (2 * a.tq);
    const int64_t plan_bytes = (bound * 8 + 255) & ~int64_t(255);
    if (!host_plan && args.max_q_len > 1 && args.workspace != nullptr && args.workspace_bytes >= 2 * plan_bytes + (int64_t(1) << 20) &&
        bound < (int64_t(1) << 24) && !(env != nullptr && env[0] == '0')) {
      dev_plan = reinterpret_cast<int32_t*>(static_cast<char*>(args.workspace) + ((args.workspace_bytes - 256) & ~int64_t(255)) - plan_bytes);
      a.work_items = dev_plan;
      n_list = bound;
    }
  }
  const int64_t plan_tail = dev_plan != nullptr ? ((static_cast<int64_t>(n_list) * 8 + 255) & ~int64_t(255)) : 0;  // workspace bytes the list takes
  const int64_t base_ctas = a.work_items != nullptr ? n_list * args.n_kv_heads : static_cast<int64_t>(n_pairs) * args.n_kv_heads * args.n_seqs;
  const int max_kv_tiles = (args.max_kv_len + kP2TileN - 1) / kP2TileN;
  int n_splits = 1;
  int direct_limit = 0;  // 0: whatever fits one split
  if (args.qk_work_hint > 0 && a.work_items != nullptr) {
    // CTA-steps: every work item walks its visible keys in 64-key steps, once per KV head
    const double total_steps = static_cast<double>(args.qk_work_hint) / kP2TileN * args.n_kv_heads;
    // Items are claimed heaviest first and just in time, so a launch is balanced unless one item is much longer than an SM's
    // share of the work: the chunk is 1.5 shares.  (Measured on B200, BASELINE config 3 with just-in-time claiming: the
    // chunked-prefill rows alone - longest item 1.44 shares - run 0.1275 ms unsplit against 0.131 / 0.132 / 0.129 ms with
    // 2 / 3 / 4 chunks, the mixed batch 0.153 ms unsplit against 0.175 / 0.193 with 2 / 3: partial traffic, the merge and the
    // extra item boundaries cost more than the last few per cent of balance.)
    // Round 2: only the HEAVY tiles are cut.  Tiles of up to one share stay whole work items (`direct_limit`), longer ones are cut
    // into pieces of about half a share, so the tail of the launch is made of half-share pieces instead of whole long items
    // (config-3 chunked prefill: 60 items of 128 steps at 94 steps per SM -> 180 pieces of 43).
    const double share = total_steps / kSms;
    double trig = 1.15;
    if (const char* env = tuning_env("HI_PAIR_HEAVY_TRIG")) trig = atof(env);
    if (max_kv_tiles > trig * share && !(tuning_env("HI_PAIR_HEAVY_SPLIT") && tuning_env("HI_PAIR_HEAVY_SPLIT")[0] == '0')) {
      double whole_frac = 1.0, piece_frac = 0.55;  // tuning overrides: HI_PAIR_WHOLE_FRAC, HI_PAIR_PIECE_FRAC
      if (const char* env = tuning_env("HI_PAIR_WHOLE_FRAC")) whole_frac = atof(env);
      if (const char* env = tuning_env("HI_PAIR_PIECE_FRAC")) piece_frac = atof(env);
      direct_limit = static_cast<int>(whole_frac * share) > kMinTilesPerSplit ? static_cast<int>(whole_frac * share) : kMinTilesPerSplit;
      double piece = piece_frac * share;
      if (piece < kMinTilesPerSplit) piece = kMinTilesPerSplit;
      const int pieces = static_cast<int>((max_kv_tiles + piece - 1) / piece);
      if (pieces > 1) n_splits = pieces;
    } else {
      int chunk = static_cast<int>(1.5 * total_steps / kSms) + 1;
      if (chunk < kMinTilesPerSplit) chunk = kMinTilesPerSplit;
      if (chunk < max_kv_tiles) n_splits = (max_kv_tiles + chunk - 1) / chunk;
    }
  } else if (base_ctas < 2 * kSms) {
    n_splits = static_cast<int>((2 * kSms + base_ctas - 1) / base_ctas);
    const int max_splits = (max_kv_tiles + kMinTilesPerSplit - 1) / kMinTilesPerSplit;
    if (n_splits > max_splits) n_splits = max_splits;
  }
  if (const char* env = tuning_env("HI_TC_SPLITS")) n_splits = atoi(env);  // tuning / test override
  if (n_splits < 1) n_splits = 1;
  const bool var_dim = args.head_dim != kP2D;
  if (var_dim) n_splits = 1;  // the fp32 partials and their merge are laid out for head_dim 128; other head dims run unsplit
  n_splits = cap_splits(n_splits, args.n_tokens, args.n_qo_heads, kP2D);
  if (n_splits > 1) {
    const int64_t need = partial_bytes_per_split(args.n_tokens, args.n_qo_heads, kP2D) * n_splits + plan_tail + kWorkspaceTailBytes;
    if (args.workspace == nullptr || need > args.workspace_bytes) {
      set_error("paged_attention: workspace of %lld bytes is smaller than the %lld needed for %d KV splits (see hi_attention_workspace_bytes)",
                (long long)args.workspace_bytes, (long long)need, n_splits);
      return HI_ERR_WORKSPACE;
    }
  }
  a.tiles_per_split = (max_kv_tiles + n_splits - 1) / n_splits;
  a.n_splits = (max_kv_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  a.direct_limit = (a.n_splits > 1 && direct_limit > a.tiles_per_split) ? direct_limit : a.tiles_per_split;
  if (a.n_splits > 1) {
    const int64_t entries = static_cast<int64_t>(args.n_tokens) * args.n_qo_heads * a.n_splits;
    a.part_o = static_cast<float*>(args.workspace);
    a.part_ml = a.part_o + entries * kP2D;
  }
  if (base_ctas * a.n_splits > 0x7fffffff) {
    set_error("paged_attention: %lld work items exceed the int32 range", static_cast<long long>(base_ctas * a.n_splits));
    return HI_ERR_INVALID_ARGUMENT;
  }
  a.max_pairs = n_pairs;
  a.n_items = static_cast<int>(base_ctas * a.n_splits);
  // dynamic work distribution: a counter in the last 256 bytes of the workspace, zeroed in stream order before the launch
  a.work_counter = nullptr;
  {
    const int64_t partial_bytes = a.n_splits > 1 ? static_cast<int64_t>(args.n_tokens) * args.n_qo_heads * a.n_splits * (kP2D + 2) * 4 : 0;
    const char* env = tuning_env("HI_PAIR_STATIC");  // tuning / test override: static boustrophedon assignment
    if (args.workspace != nullptr && args.workspace_bytes >= partial_bytes + plan_tail + 512 && !(env != nullptr && env[0] == '1')) {
      a.work_counter = reinterpret_cast<unsigned int*>(static_cast<char*>(args.workspace) + ((args.workspace_bytes - 256) & ~int64_t(255)));
      if (dev_plan == nullptr) HI_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned int), stream));  // else the plan kernel zeroes it
    }
  }
  if (dev_plan != nullptr) {
    HI_CUDA(launch_pdl(p2_plan_kernel, dim3(1), dim3(kPlanThreads), 0, stream, args.q_cu_seq_lens, args.kv_cu_seq_lens, args.n_seqs, 2 * a.tq,
                       args.max_kv_len, static_cast<int>(n_list), dev_plan, a.work_counter));
    note_launch();
    HI_CUDA(cudaGetLastError());
  }

  CUtensorMap mq, mk, mv, mo;
  int rc = make_map_d(&mq, args.dtype, args.q, args.n_tokens, args.n_qo_heads, args.head_dim, args.q_row_stride, a.group, a.tq);
  if (rc != HI_OK) return rc;
  rc = make_map_d(&mo, args.dtype, args.out, args.n_tokens, args.n_qo_heads, args.head_dim, args.out_row_stride, a.group, a.tq);
  if (rc != HI_OK) return rc;
  const int64_t n_slots = args.n_blocks * args.block_size;
  rc = pool_map_d(&mk, args.dtype, args.key_cache, n_slots, args.n_kv_heads, args.head_dim, args.block_size);
  if (rc != HI_OK) return rc;
  rc = pool_map_d(&mv, args.dtype, args.value_cache, n_slots, args.n_kv_heads, args.head_dim, args.block_size);
  if (rc != HI_OK) return rc;
  if (const char* env = tuning_env("HI_PAIR_DEBUG")) a.debug = atoi(env);
  if (var_dim) {
    // MODE 2: the tensor maps carry the real head_dim (TMA zero-fills up to the 64-dim box), the MMAs walk head_dim / 16 k-steps
    a.head_dim = args.head_dim;
    a.causal = 1;
    a.n_halves = args.head_dim > 64 ? 2 : 1;
    a.n_kk = args.head_dim / 16;
    a.idesc_pv = ptx::make_idesc_f16(args.dtype == HI_BF16, false, true, kP2TileM, args.head_dim);
    a.use_tma_store = 0;
    // (With half the head dim the products need half the tensor-pipe cycles while the softmax still needs one exponential per
    // score, so the softmax warps bound the kernel: 8k prefill at head_dim 64 runs 0.331 ms against 0.393 ms at 128.  Moving
    // 1/4 or 1/2 of the exponentials from MUFU.EX2 to an FMA-pipe polynomial was measured on B200 and does not help - 0.331 /
    // 0.331 / 0.366 ms: the warps are short of issue slots, not of MUFU throughput.)
    return args.dtype == HI_BF16 ? launch_pair_t<__nv_bfloat16, 0, 2>(args.device, a, mq, mk, mv, mo, stream)
                                 : launch_pair_t<__half, 0, 2>(args.device, a, mq, mk, mv, mo, stream);
  }

  // decode rows of a mixed batch are walked by both tiles (the exchange rows of kP2Xchg floats fit tile 1's 16 KiB staging buffer)
  a.interleave_decode = a.group * kP2Xchg * 4 <= kP2QHalf ? 1 : 0;
  if (const char* env = tuning_env("HI_PAIR_INTERLEAVE")) a.interleave_decode = a.interleave_decode && atoi(env) != 0;  // A/B switch
  a.use_tma_store = 1;
#ifdef HI_P2_NO_STAGING
  a.use_tma_store = 0;
#endif
  if (const char* env = tuning_env("HI_PAIR_TMA_STORE")) a.use_tma_store = atoi(env) != 0;  // A/B switch
  int poly = 0;  // exponentials per 4 moved from MUFU to the FMA pipes (measured: no gain while the softmax warps have idle issue slots)
  if (const char* env = tuning_env("HI_PAIR_POLY")) poly = atoi(env);  // tuning override
  if (args.dtype == HI_BF16) {
    rc = poly <= 0 ? launch_pair_t<__nv_bfloat16, 0>(args.device, a, mq, mk, mv, mo, stream)
       : poly == 1 ? launch_pair_t<__nv_bfloat16, 1>(args.device, a, mq, mk, mv, mo, stream)
                   : launch_pair_t<__nv_bfloat16, 2>(args.device, a, mq, mk, mv, mo, stream);
  } else {
    rc = poly <= 0 ? launch_pair_t<__half, 0>(args.device, a, mq, mk, mv, mo, stream)
       : poly == 1 ? launch_pair_t<__half, 1>(args.device, a, mq, mk, mv, mo, stream)
                   : launch_pair_t<__half, 2>(args.device, a, mq, mk, mv, mo, stream);
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
  m.chunk_tiles = a.tiles_per_split * (kP2TileN / 16);
  m.part_o = a.part_o;
  m.part_ml = a.part_ml;
  m.direct_tile_tokens = a.tq;
  m.direct_tiles = a.direct_limit * (kP2TileN / 16);
  return launch_merge_partials(m, args.dtype, kP2D, stream);
}

// ---- un-paged varlen attention (hi_varlen_attention): the vision-encoder form of mha_varlen_fwd ------------------------------
bool varlen_pair_supported(const HiVarlenArgs& v) {
  const int group = v.n_kv_heads > 0 ? v.n_qo_heads / v.n_kv_heads : 0;
  return (v.dtype == HI_F16 || v.dtype == HI_BF16) && v.head_dim >= 8 && v.head_dim <= kP2D && (v.head_dim % 8) == 0 &&
         group >= 1 && group <= kP2TileM && (v.q_row_stride % 8) == 0 && (v.k_row_stride % 8) == 0 &&
         (v.v_row_stride % 8) == 0 && (v.out_row_stride % 8) == 0 && aligned_to(v.q, 16) && aligned_to(v.k, 16) &&
         aligned_to(v.v, 16) && aligned_to(v.out, 16);
}

int launch_varlen_pair(const HiVarlenArgs& v, cudaStream_t stream) {
  if (!varlen_pair_supported(v)) {
    set_error("varlen_attention: needs fp16/bf16, head_dim a multiple of 8 up to 128 and 16-byte aligned rows (got dtype %d head_dim %d)", v.dtype, v.head_dim);
    return HI_ERR_UNSUPPORTED;
  }
  P2Args a{};
  a.out = v.out;
  a.out_row_stride = v.out_row_stride;
  a.q_cu = v.cu_seqlens_q;
  a.kv_cu = v.cu_seqlens_k;
  a.n_qo_heads = v.n_qo_heads;
  a.n_kv_heads = v.n_kv_heads;
  a.group = v.n_qo_heads / v.n_kv_heads;
  a.block_size = kP2TileN;  // a "page" is one 64-key step of consecutive rows
  a.tq = kP2TileM / a.group;
  a.scale_log2 = v.softmax_scale * 1.4426950408889634f;
  a.head_dim = v.head_dim;
  a.causal = v.causal ? 1 : 0;
  a.n_halves = v.head_dim > 64 ? 2 : 1;
  a.n_kk = (v.head_dim + 15) / 16;
  a.idesc_pv = ptx::make_idesc_f16(v.dtype == HI_BF16, false, true, kP2TileM, a.n_kk * 16);
  a.n_splits = 1;  // sequences of a vision batch are plentiful and short: no split-KV
  a.tiles_per_split = (v.max_kv_len + kP2TileN - 1) / kP2TileN;
  a.direct_limit = a.tiles_per_split;
  a.max_pairs = (v.max_q_len + 2 * a.tq - 1) / (2 * a.tq);
  const int64_t n_items = static_cast<int64_t>(a.max_pairs) * v.n_kv_heads * v.n_seqs;
  if (n_items > 0x7fffffff) {
    set_error("varlen_attention: %lld work items exceed the int32 range", (long long)n_items);
    return HI_ERR_INVALID_ARGUMENT;
  }
  a.n_items = static_cast<int>(n_items);
  a.work_counter = nullptr;
  if (v.workspace != nullptr && v.workspace_bytes >= 512) {
    a.work_counter = reinterpret_cast<unsigned int*>(static_cast<char*>(v.workspace) + ((v.workspace_bytes - 256) & ~int64_t(255)));
    HI_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned int), stream));
  }
  CUtensorMap mq, mk, mv, mo;
  int rc = make_map_d(&mq, v.dtype, v.q, v.n_q_tokens, v.n_qo_heads, v.head_dim, v.q_row_stride, a.group, a.tq);
  if (rc != HI_OK) return rc;
  rc = make_map_d(&mo, v.dtype, v.out, v.n_q_tokens, v.n_qo_heads, v.head_dim, v.out_row_stride, a.group, a.tq);
  if (rc != HI_OK) return rc;
  a.use_tma_store = v.head_dim == kP2D ? 1 : 0;
#ifdef HI_P2_NO_STAGING
  a.use_tma_store = 0;
#endif
  if (const char* env = tuning_env("HI_PAIR_TMA_STORE")) a.use_tma_store = a.use_tma_store && atoi(env) != 0;
  rc = make_map_d(&mk, v.dtype, v.k, v.n_k_tokens, v.n_kv_heads, v.head_dim, v.k_row_stride, 1, kP2TileN);
  if (rc != HI_OK) return rc;
  rc = make_map_d(&mv, v.dtype, v.v, v.n_k_tokens, v.n_kv_heads, v.head_dim, v.v_row_stride, 1, kP2TileN);
  if (rc != HI_OK) return rc;
  if (const char* env = tuning_env("HI_PAIR_DEBUG")) a.debug = atoi(env);
  return v.dtype == HI_BF16 ? launch_pair_t<__nv_bfloat16, 0, 1>(v.device, a, mq, mk, mv, mo, stream)
                            : launch_pair_t<__half, 0, 1>(v.device, a, mq, mk, mv, mo, stream);
}

}  // namespace hi

#ifdef HI_MBAR_DEBUG
// Debug builds only: copies (and clears) the first mbarrier timeout recorded by the pair kernel.
extern "C" int hi_debug_mbar_timeout(unsigned int out[64]) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, hi::ptx::g_mbar_debug, 64 * sizeof(unsigned int)) != cudaSuccess) return -1;
  unsigned int zero[64] = {};
  cudaMemcpyToSymbol(hi::ptx::g_mbar_debug, zero, sizeof(zero));
  return 0;
}
#endif

#ifdef HI_PAIR_TRACE
// Dev builds only: copies the timeline CTA 0 recorded during the last pair-kernel launch (see tools/pair_trace.py).
extern "C" int hi_debug_pair_trace(unsigned long long* records /* [kTraceRoles][kTraceCap] */, unsigned int* counts /* [kTraceRoles] */) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(records, hi::g_pair_trace, sizeof(unsigned long long) * hi::kTraceRoles * hi::kTraceCap) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(counts, hi::g_pair_trace_n, sizeof(unsigned int) * hi::kTraceRoles) != cudaSuccess) return -1;
  return 0;
}
// Dev builds only: per-CTA {start ns, end ns, items, steps} of the last pair-kernel launch (load balance of the item walk).
extern "C" int hi_debug_pair_cta_stats(unsigned long long* stats /* [1024][4] */) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(stats, hi::g_pair_cta_stats, sizeof(unsigned long long) * 1024 * 4) == cudaSuccess ? 0 : -1;
}
#endif
