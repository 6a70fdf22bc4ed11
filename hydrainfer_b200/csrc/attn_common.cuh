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
};

// One merged row = LSE-weighted sum of its valid chunks; chunk c covers 16-token tiles [c*chunk_tiles, (c+1)*chunk_tiles).
int launch_merge_partials(const SimtArgs& a, int dtype, int head_dim, cudaStream_t stream);

}  // namespace hi
