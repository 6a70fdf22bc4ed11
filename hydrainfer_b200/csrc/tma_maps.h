// TMA tensor-map construction shared by the tcgen05 kernels (implemented in attn_tc.cu).
#pragma once

#include <cuda.h>

#include <cstdint>

namespace hi {

// 3-D map over [rows, heads, 128] 16-bit elements (row stride in elements); box = {64 dims, box_heads, box_rows},
// 128-byte swizzle.  Returns a HiStatus.
int make_map(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t heads, int64_t row_stride_elems,
             int box_heads, int box_rows);

// Same for rows of `head_dim` (a multiple of 8, <= 128) elements per head: the box is still 64 dims wide and TMA zero-fills
// whatever lies beyond head_dim.
int make_map_d(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t heads, int head_dim, int64_t row_stride_elems,
               int box_heads, int box_rows);

// Cached map of a paged pool [n_slots, heads, 128]; box = one page of one head and one 64-dim half.
int pool_map(CUtensorMap* out, int dtype, const void* base, int64_t n_slots, int heads, int block_size);

// Same for pools whose heads are `head_dim` (a multiple of 8, <= 128) elements wide.
int pool_map_d(CUtensorMap* out, int dtype, const void* base, int64_t n_slots, int heads, int head_dim, int block_size);

}  // namespace hi
