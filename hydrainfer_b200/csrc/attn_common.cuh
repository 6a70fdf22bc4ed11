// Argument block shared by the split-KV CUDA-core kernels and the split-merge kernel (also used by the tile kernel's
// split-KV mode to launch the merge).
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

namespace hi {

struct SimtArgs {
  const void* q;
  void* out;
  const void* kc;
  const void* vc;
  int64_t q_row_stride, out_row_stride;  // elements
  int64_t tok_stride;                    // elements between consecutive slots of the cache: Hkv * D
  const int32_t* q_cu;
  const int32_t* kv_cu;
  const int32_t* block_tables;
  const int32_t* cu_blocks;
  int n_seqs, n_tokens, n_qo_heads, n_kv_heads, group, block_size;
  int chunk_tiles;  // 16-token tiles per chunk
  int n_chunks;
  float scale_log2;  // softmax_scale * log2(e)
  float* part_o;     // [n_tokens * Hq * n_chunks][D]
  float* part_ml;    // [n_tokens * Hq * n_chunks][2]
  // Merge only: > 0 when the producer handled query rows in tiles of this many tokens per sequence and wrote a tile whose
  // LAST row needs a single chunk straight to `out` (normalised); the merge leaves those rows alone.
  int direct_tile_tokens;
  int direct_tiles;  // ... with at most this many 16-token tiles of visible keys (0: chunk_tiles, i.e. "fits the first chunk")
  // mha_varlen_fwd's score options (flash_api.cpp:93-111, src/mask.h:54-62, 158-195, flash_fwd_kernel.h:244-245, 283); read by the
  // OPT instances of paged_attn_simt_kernel only
  float softcap_log2;      // softcap * log2(e); 0 = off.  score = softcap * tanh(q.k * scale / softcap)
  float inv_softcap_log2;  // 1 / softcap_log2
  int window_left;         // keys below i_abs - window_left are masked; < 0 = unlimited
  int window_right;        // keys above i_abs + window_right are masked; < 0 = unlimited (0 = causal)
  const float* alibi_slopes;     // [n_qo_heads] or [n_seqs][n_qo_heads] fp32, NULL = off.  score -= slope * |i_abs - j|
  int64_t alibi_batch_stride;    // 0 for the per-head form
};

// One merged row = LSE-weighted sum of its valid chunks; chunk c covers 16-token tiles [c*chunk_tiles, (c+1)*chunk_tiles).
int launch_merge_partials(const SimtArgs& a, int dtype, int head_dim, cudaStream_t stream);

}  // namespace hi
