/*
 * hi_b200.h — C ABI of the B200-native paged-KV attention hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Each entry point names the
 * reference interface it replaces (paths relative to the hydrainfer repository).  All functions
 *   - return 0 on success or a negative HiStatus; hi_last_error() gives a human-readable message
 *     for the calling thread (the reference instead aborts the process: glog CHECK in
 *     csrc/kernel/kv_cache_kernels/kv_cache_kernels.cu:67-68, printf+exit in
 *     csrc/data_transfer/block_migration.cpp:9-15);
 *   - launch asynchronously on the cudaStream_t passed as `stream` (an opaque pointer here so the header
 *     needs no CUDA include); the reference uses at::cuda::getCurrentCUDAStream()
 *     (kv_cache_kernels.cu:83, flash_api.cpp:353, block_migration.cpp:201) — the host shim passes that;
 *   - never allocate device memory: scratch comes from the caller-owned workspace.
 *
 * Device pointers are marked [dev]; everything else is host memory.
 */
#ifndef HI_B200_H_
#define HI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HI_B200_ABI_VERSION 6

typedef enum HiStatus {
  HI_OK = 0,
  HI_ERR_INVALID_ARGUMENT = -1, /* layout / shape precondition violated */
  HI_ERR_UNSUPPORTED = -2,      /* dtype / head_dim / block size outside what the kernels cover */
  HI_ERR_CUDA = -3,             /* a CUDA runtime or driver call failed */
  HI_ERR_WORKSPACE = -4,        /* workspace too small; see hi_attention_workspace_bytes */
  HI_ERR_PEER_UNSUPPORTED = -5  /* peer access to the exporting GPU is not possible */
} HiStatus;

typedef enum HiDtype { HI_F32 = 0, HI_F16 = 1, HI_BF16 = 2 } HiDtype;

/* Which attention kernel hi_paged_attention should run. AUTO is the production choice. */
typedef enum HiAttnPath {
  HI_ATTN_AUTO = 0,
  HI_ATTN_SIMT = 1,   /* split-KV CUDA-core kernel (decode rows; also the generic any-shape path) */
  HI_ATTN_TCGEN05 = 2, /* tcgen05/TMEM tile kernel (prefill, chunked prefill, GQA-packed decode) */
  HI_ATTN_TCGEN05_DECODE = 3, /* tcgen05 swapped-operand kernel: one query row x one KV head per CTA (grouped decode) */
  HI_ATTN_TCGEN05_PAIR = 4 /* tcgen05 pair-tile kernel: two ping-ponged 128-row query tiles per CTA (prefill, chunked prefill) */
} HiAttnPath;

const char* hi_last_error(void);
int hi_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * KV append — replaces set_kv_cache (csrc/kernel/kv_cache_kernels/kv_cache_kernels.cu:60-95) and
 * set_image_cache (csrc/kernel/cache_kernels/cache_kernels.cu:55-83).
 *
 *   cache[slot / block_size, slot % block_size, :, :] = rows[t, :, :]      (bit-exact copy)
 *
 * Caches are contiguous [n_blocks, block_size, n_heads, head_dim], so the destination of token t is
 * the `row_elems`-long run starting at slot_ids[t] * row_elems.  Source rows are contiguous over
 * (n_heads, head_dim) but may have a larger row stride (a slice of a fused qkv projection,
 * kv_cache_kernels.cu:75-76).  hi_set_kv_cache scatters K and V in ONE launch.
 * ------------------------------------------------------------------------------------------- */
int hi_set_kv_cache(const int32_t* slot_ids /*[dev] [n_tokens]*/,
                    const void* keys /*[dev]*/, const void* values /*[dev]*/,
                    void* key_cache /*[dev]*/, void* value_cache /*[dev]*/,
                    int64_t n_tokens, int64_t row_elems /* n_kv_heads*head_dim */,
                    int64_t key_row_stride /*elements*/, int64_t value_row_stride /*elements*/,
                    int dtype /*HiDtype*/, int device, void* stream);

int hi_set_image_cache(const int32_t* slot_ids /*[dev] [n_tokens]*/,
                       const void* image_tokens /*[dev]*/, void* image_cache /*[dev]*/,
                       int64_t n_tokens, int64_t row_elems /* n_heads*head_dim */,
                       int64_t token_row_stride /*elements*/,
                       int dtype, int device, void* stream);

/* Image-embedding read-back — replaces the advanced-indexing gather `image_token_cache[slot_ids, :]` of
 * LanguageModelParametersBuilder.add (hydrainfer/engine/parameters_builder.py:48-55), the consumer side of
 * set_image_cache:   out[t, :] = cache[slot_ids[t], :]   (bit-exact copy; cache viewed as [n_slots, row_elems]). */
int hi_get_image_cache(const int32_t* slot_ids /*[dev] [n_tokens]*/, const void* image_cache /*[dev]*/,
                       void* out /*[dev] [n_tokens, row_elems], row stride out_row_stride*/,
                       int64_t n_tokens, int64_t row_elems, int64_t out_row_stride /*elements*/,
                       int dtype, int device, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rotary embedding fused with the KV append — replaces apply_rotary_pos_emb
 * (csrc/kernel/position_embedding/rope.cu:90-117, bound in position_embedding_pybind.cpp) and, when slot_ids is not
 * NULL, also the set_kv_cache launch that ROPECausalGroupedQueryPageAttention.forward issues right after it
 * (hydrainfer/model/model_forward.py:81-83 -> hydrainfer/layer/causal_attention.py:402).
 *
 *   pair (x, y) of a head, c/s = cos/sin of the token's position:  x' = x*c - y*s,  y' = x*s + y*c
 *   interleaved: (x, y) = (head[2i], head[2i+1]);  otherwise (head[i], head[i + rotary_dim/2])   (rope.cu:19-29)
 *   dims >= rotary_dim pass through.  q is rotated in place.  With slot_ids: the rotated k and v go to
 *   key_cache / value_cache[slot_ids[t]] (geometry as hi_set_kv_cache) and k itself is rewritten only if
 *   write_back_k; without slot_ids k is rotated in place (plain apply_rotary_pos_emb).
 *
 * Rounding is that of the reference's torch handler (TorchRotaryEmbeddingHandler, hydrainfer/layer/rotary_embedding.py:44-83):
 * every product and the final sum round separately, to the element type when the table has the element type, to fp32 when
 * the table is fp32 (torch promotes).  Results are bit-identical to that path.  (The reference's compiled kernel computes
 * `x*c - y*s` in native half, which nvcc contracts into one FMA - one rounding fewer - and has no bf16 instance,
 * csrc/kernel/dispatch.h:12-29; tests/test_gpu_reference_native.py bounds the difference.)
 * ------------------------------------------------------------------------------------------- */
typedef struct HiRopeArgs {
  void* q;                   /* [dev] [n_tokens, n_qo_heads, head_dim], head stride == head_dim; rotated in place */
  void* k;                   /* [dev] [n_tokens, n_kv_heads, head_dim] */
  const void* v;             /* [dev] [n_tokens, n_kv_heads, head_dim]; only read when slot_ids != NULL */
  int64_t q_row_stride, k_row_stride, v_row_stride; /* elements between consecutive tokens */
  const void* positions;     /* [dev] [n_tokens] int32 (rope.cu:107) or int64 */
  const void* cos_sin;       /* [dev] [max_positions, 2, rotary_dim/2]: cos then sin per position (rotary_embedding.py:113-115) */
  const int32_t* slot_ids;   /* [dev] [n_tokens] physical cache slots, or NULL for rotation only */
  void* key_cache;           /* [dev] [n_blocks, block_size, n_kv_heads, head_dim] contiguous */
  void* value_cache;         /* [dev] same geometry */
  int64_t n_tokens;
  int32_t n_qo_heads, n_kv_heads, head_dim, rotary_dim;
  int32_t dtype;             /* HiDtype of q/k/v/caches */
  int32_t cos_sin_dtype;     /* HiDtype of the table: == dtype or HI_F32 */
  int32_t positions_int64;   /* 0: int32 positions, 1: int64 */
  int32_t interleaved;
  int32_t write_back_k;      /* with slot_ids: also rewrite k in place (the reference always does) */
  int32_t force_scalar;      /* tests: take the any-shape scalar kernel even when the vector kernel applies */
  int32_t device;
  int32_t reserved;
} HiRopeArgs;

int hi_rope_append(const HiRopeArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Paged causal GQA attention — replaces mha_varlen_fwd (csrc/kernel/flash_attn/flash_api.cpp:216-355)
 * as called by FlashAttentionCausalGroupedQueryPageAttentionHandler
 * (hydrainfer/layer/causal_attention.py:262-294), the flashinfer run() of :235-249, and computes what
 * TorchCausalGroupedQueryPageAttentionHandler (:307-374) computes:
 *
 *   for sequence b, query row i (0-based among its q_b new tokens), key j of its L_b cached tokens:
 *       visible(i, j)  <=>  j <= i + L_b - q_b            (bottom-right aligned causal mask, :339-342)
 *       o[i, h] = softmax_j(scale * q[i, h] . k[j, h / group]) @ v[:, h / group]
 *
 * KV has already been appended (causal_attention.py:402-403); L_b includes the new tokens.
 * The metadata arrays are the int32 device tensors of AttentionParameters (:31-68):
 * block_tables is the flattened CSR of per-sequence block ids, cu_blocks_lens its row pointer.
 * ------------------------------------------------------------------------------------------- */
typedef struct HiAttnArgs {
  /* tensors */
  const void* q;             /* [dev] [n_tokens, n_qo_heads, head_dim], head stride == head_dim */
  void* out;                 /* [dev] same shape; written in place like mha_varlen_fwd's `out` */
  const void* key_cache;     /* [dev] [n_blocks, block_size, n_kv_heads, head_dim] contiguous */
  const void* value_cache;   /* [dev] same geometry */
  int64_t q_row_stride;      /* elements between consecutive tokens of q */
  int64_t out_row_stride;    /* elements between consecutive tokens of out */
  /* metadata (AttentionParameters) */
  const int32_t* q_cu_seq_lens;  /* [dev] [n_seqs + 1] */
  const int32_t* kv_cu_seq_lens; /* [dev] [n_seqs + 1] */
  const int32_t* block_tables;   /* [dev] [sum ceil(L_b / block_size)] */
  const int32_t* cu_blocks_lens; /* [dev] [n_seqs + 1] */
  int32_t n_seqs;
  int32_t n_tokens;
  int32_t max_q_len;         /* AttentionParameters.q_max_seq_len */
  int32_t max_kv_len;        /* AttentionParameters.kv_max_seq_len */
  /* geometry */
  int32_t n_qo_heads, n_kv_heads, head_dim, block_size;
  int64_t n_blocks;          /* pool blocks (bounds the TMA tensor map) */
  int32_t dtype;             /* HiDtype of q/out/caches */
  float softmax_scale;       /* 1/sqrt(head_dim) in the reference */
  /* scratch + control */
  void* workspace;           /* [dev] >= hi_attention_workspace_bytes(...) bytes, 256-B aligned */
  int64_t workspace_bytes;
  int32_t path;              /* HiAttnPath */
  int32_t device;
  int32_t kv_blocks_hint;    /* entries of block_tables (sum of ceil(L_b / block_size)); 0 = unknown.  Only steers the
                                split-KV heuristics (ragged batches are split finer); never affects results. */
  int32_t reserved[3];
  /* ABI 2: optional host-side plan.  The host builds the metadata from Python ints (AttentionParametersBuilder), so it
   * can also enumerate the query tiles of the batch, sorted by cost, the way flashinfer's plan() does for the reference
   * (causal_attention.py:171-195).  The plan steers work order, grid size and split-KV only; results never depend on it.
   * Tiles are runs of hi_attention_tile_tokens(n_qo_heads, n_kv_heads) consecutive query tokens of one sequence. */
  const int32_t* work_items; /* [dev] [n_work_items][2] = (sequence, tile index within the sequence), heaviest first; NULL = none:
                                batches with prefill rows then get the same list from a one-CTA kernel that sorts the tiles
                                on the device (in the tail of `workspace`; one extra launch of a few microseconds) */
  int64_t qk_work_hint;      /* sum over the work items of the number of keys the item walks, i.e. of
                                L_b - q_b + min(q_b, (tile + 1) * tile_tokens); 0 = unknown */
  int32_t n_work_items;
  int32_t work_tile_tokens;  /* tile size the plan was built for; the plan is ignored if it does not match the kernel's */
  /* ABI 6: the score options of mha_varlen_fwd (flash_api.cpp:225-232; the reference's FlashAttention-2 build has alibi, local
   * windows and softcap enabled, :93-111, :197-213).  The paged attention layer never passes them (causal_attention.py:274-291),
   * so they run on the any-shape CUDA-core kernel (head_dim 64 / 128 / 256, unsplit), not on the tcgen05 kernels.
   * With i_abs = i + kv_len - q_len (the query's position in key coordinates, src/mask.h:54-55):
   *   HI_ATTN_OPT_WINDOW   key j is visible iff i_abs - window_left <= j <= i_abs + window_right, a negative bound = unlimited
   *                        (the default without the flag is window_left = -1, window_right = 0: causal)
   *   HI_ATTN_OPT_SOFTCAP  score = softcap * tanh(q.k * softmax_scale / softcap)                        (softcap > 0)
   *   HI_ATTN_OPT_ALIBI    score -= alibi_slopes[b * alibi_batch_stride + h] * |i_abs - j|, applied after the softcap
   * A row without a visible key gets zeros. */
  int32_t options;           /* HiAttnOptions bit mask; 0 = plain causal attention */
  int32_t window_left, window_right;
  float softcap;
  const float* alibi_slopes; /* [dev] fp32, [n_qo_heads] (alibi_batch_stride 0) or [n_seqs, n_qo_heads] */
  int64_t alibi_batch_stride;
} HiAttnArgs;

typedef enum HiAttnOptions { HI_ATTN_OPT_WINDOW = 1, HI_ATTN_OPT_SOFTCAP = 2, HI_ATTN_OPT_ALIBI = 4 } HiAttnOptions;

/* Upper bound of the scratch hi_paged_attention needs for a batch with these extents (split-KV partials). */
int64_t hi_attention_workspace_bytes(int32_t n_tokens, int32_t n_qo_heads, int32_t head_dim, int32_t max_kv_len);

int hi_paged_attention(const HiAttnArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Un-paged varlen attention — the OTHER form of mha_varlen_fwd (flash_api.cpp:216-355 with block_table == None):
 * k and v are plain [n_k_tokens, n_kv_heads, head_dim] tensors and sequence b owns rows cu_seqlens_k[b] .. [b+1].
 * This is what the vision encoders call: FlashAttentionMutliHeadAttentionHandler2.forward and
 * QwenFlashAttentionMutliHeadAttentionHandler2.forward (hydrainfer/layer/multihead_attention.py:118-160, 183-211)
 * pass window (-1, -1) = NOT causal; it computes what TorchMultiHeadAttentionHandler (:48-73) and
 * QwenTorchMultiHeadAttentionHandler (:241-256) compute:
 *
 *       o[i, h] = softmax_j(scale * q[i, h] . k[j, h / group]) @ v[:, h / group]      j over the sequence's keys,
 *       restricted to j <= i + L_b - q_b when `causal`.
 *
 * tcgen05 pair-tile kernel; fp16 / bf16, head_dim any multiple of 8 up to 128 (CLIP 64, SigLIP 72, Qwen2-VL 80, 128).
 * ------------------------------------------------------------------------------------------- */
typedef struct HiVarlenArgs {
  const void* q;             /* [dev] [n_q_tokens, n_qo_heads, head_dim], head stride == head_dim */
  const void* k;             /* [dev] [n_k_tokens, n_kv_heads, head_dim] */
  const void* v;             /* [dev] [n_k_tokens, n_kv_heads, head_dim] */
  void* out;                 /* [dev] [n_q_tokens, n_qo_heads, head_dim]; written in place */
  int64_t q_row_stride, k_row_stride, v_row_stride, out_row_stride; /* elements between consecutive tokens */
  const int32_t* cu_seqlens_q; /* [dev] [n_seqs + 1] */
  const int32_t* cu_seqlens_k; /* [dev] [n_seqs + 1] */
  int32_t n_seqs, n_q_tokens, n_k_tokens;
  int32_t max_q_len, max_kv_len;
  int32_t n_qo_heads, n_kv_heads, head_dim;
  int32_t dtype;             /* HiDtype */
  int32_t causal;            /* window_size_right == 0 in the reference's call */
  float softmax_scale;
  int32_t device;
  void* workspace;           /* [dev] optional (>= 512 bytes enables dynamic work distribution) */
  int64_t workspace_bytes;
  int64_t reserved[2];
} HiVarlenArgs;

int hi_varlen_attention(const HiVarlenArgs* args, void* stream);

/* Query tokens per work item of the prefill kernel for this head geometry (0 if that kernel does not cover it). */
int32_t hi_attention_tile_tokens(int32_t n_qo_heads, int32_t n_kv_heads);

/* Number of kernels the last API call on this thread launched (bench bookkeeping). */
int hi_last_launch_count(void);

/* Measurement hooks (bench.py's roofline leg).  Events live in this library's CUDA runtime instance.
 * hi_set_kernel_timing_events(start, stop): the NEXT hi_paged_attention call on this thread records `start`
 * immediately before and `stop` immediately after its dominant kernel (the split-KV / tile kernel, excluding the
 * split-merge kernel) on the launch stream; pass NULLs to disarm.  The pair is consumed by that call. */
int hi_event_create(void** event_out);
int hi_event_destroy(void* event);
int hi_event_record(void* event, void* stream);
int hi_event_elapsed_ms(void* start, void* stop, float* ms_out); /* synchronizes on `stop` */
int hi_set_kernel_timing_events(void* start, void* stop);

/* ---------------------------------------------------------------------------------------------
 * KV-page migration — replaces csrc/data_transfer/block_migration.cpp.
 *
 * Pools are contiguous [n_layers, n_tokens, n_blocks, block_size, n_heads, head_size]
 * (hydrainfer/memory/token_cache_manger.py:65).  For every (layer, kv, i) the run of
 * run_bytes = block_size*n_heads*head_size*itemsize bytes of src block src_blocks[i] is copied to dst block
 * dst_blocks[i] (block_migration.cpp:222-244); pools may differ in n_blocks only.
 * The reference issues n_layers*n_tokens*n cudaMemcpyAsync calls; this is ONE launch: when either pool is
 * memory of another GPU (a peer pointer or a CUDA-IPC mapping) the TMA engine moves the runs as 1-D bulk copies
 * through shared memory, issued by one warp per SM (the receiver's SMs stay with its decode step); copies inside
 * one GPU use a load / store gather kernel.
 * ------------------------------------------------------------------------------------------- */
typedef struct HiPoolGeom {
  int64_t n_layers, n_tokens, n_blocks, run_bytes;
} HiPoolGeom;

int hi_migrate_blocks(const int32_t* src_blocks /*[dev] [n]*/, const int32_t* dst_blocks /*[dev] [n]*/,
                      int64_t n, const void* src_pool /*[dev] local or peer-mapped*/, void* dst_pool /*[dev]*/,
                      HiPoolGeom src, HiPoolGeom dst, int device, void* stream);

/* The same copy restricted to layers [layer_begin, layer_end): a sender that has finished layer l of a prefill can ship
 * that layer's pages while layer l + 1 is still computing (the reference moves whole requests after the last layer,
 * hydrainfer/engine/executor.py migrate path).  The launch may equally run on the SOURCE GPU with dst_pool peer-mapped
 * (push): the kernel only needs both pointers to be addressable from `device`. */
int hi_migrate_blocks_layers(const int32_t* src_blocks /*[dev] [n]*/, const int32_t* dst_blocks /*[dev] [n]*/,
                             int64_t n, const void* src_pool /*[dev]*/, void* dst_pool /*[dev]*/,
                             HiPoolGeom src, HiPoolGeom dst, int64_t layer_begin, int64_t layer_end,
                             int device, void* stream);

/* The same copy with the block tables in HOST memory (what migrate_blocks of block_migration.cpp:194-199 receives: two
 * std::vector<int64_t>): requests of up to hi_migrate_inline_table_blocks() blocks carry their tables inside the kernel
 * parameters, so the launch needs no staging buffer and no host-to-device copy in front of it.  Larger requests: upload the
 * tables and call hi_migrate_blocks_layers. */
int hi_migrate_blocks_host_tables(const int32_t* src_blocks_host /*[n]*/, const int32_t* dst_blocks_host /*[n]*/, int64_t n,
                                  const void* src_pool /*[dev]*/, void* dst_pool /*[dev]*/, HiPoolGeom src, HiPoolGeom dst,
                                  int64_t layer_begin, int64_t layer_end, int device, void* stream);
int hi_migrate_inline_table_blocks(void);

/* Caps the grid of the migration kernels (process-wide; 0 = default, 8 CTAs per SM).  A receiver that decodes while it pulls
 * pages (hydrainfer/cluster/epdnode.py:362-447 issues the pull on a side stream during the step) trades transfer rate for
 * decode throughput with it; bench.py's `migrate_under_decode` extra measures both sides. */
int hi_migrate_set_max_ctas(int max_ctas);

/* get_ipc_mem_handle (block_migration.cpp:55-59).  handle_out receives the 64 bytes of
 * cudaIpcMemHandle_t of the ALLOCATION containing ptr; *offset_out the byte offset of ptr inside it
 * (the reference silently assumes 0). */
int hi_ipc_get_handle(const void* ptr /*[dev]*/, uint8_t handle_out[64], int64_t* offset_out, int device);

/* register_ipc_mem_handle (block_migration.cpp:69-80) + the per-call cudaIpcOpenMemHandle of :213-215,
 * cached per process: the same handle is mapped once.  *ptr_out = mapped base + offset. */
int hi_ipc_open_handle(const uint8_t handle[64], int64_t offset, int device, void** ptr_out);

/* Unmap every cached peer mapping (the reference never closes them). */
int hi_ipc_close_all(void);

/* cudaDeviceEnablePeerAccess(peer_device) on `device`, idempotent.  Needed only when source and destination pools
 * live in the SAME process on different GPUs (IPC mappings enable peer access lazily by themselves). */
int hi_enable_peer_access(int device, int peer_device);

/* cudaMemcpyPeerAsync of one contiguous range — the NVLink roofline probe used by bench/tests. */
int hi_peer_copy(void* dst, int dst_device, const void* src, int src_device, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HI_B200_H_ */
