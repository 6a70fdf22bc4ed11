// hi_paged_attention: argument validation and kernel selection, plus the error / launch-count plumbing
// shared by every entry point of include/hi_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

static_assert(sizeof(HiAttnArgs) == 224, "HiAttnArgs layout is part of the ABI (ctypes mirror in hydrainfer_b200/_lib.py)");
static_assert(sizeof(HiPoolGeom) == 32, "HiPoolGeom layout is part of the ABI");
static_assert(sizeof(HiRopeArgs) == 144, "HiRopeArgs layout is part of the ABI");
static_assert(sizeof(HiVarlenArgs) == 160, "HiVarlenArgs layout is part of the ABI");

namespace hi {

static thread_local char g_error[512] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
static thread_local cudaEvent_t g_ev_start = nullptr;
static thread_local cudaEvent_t g_ev_stop = nullptr;
void timing_mark_start(cudaStream_t stream) {
  if (g_ev_start != nullptr) {
    cudaEventRecord(g_ev_start, stream);
    g_ev_start = nullptr;
  }
}
void timing_mark_stop(cudaStream_t stream) {
  if (g_ev_stop != nullptr) {
    cudaEventRecord(g_ev_stop, stream);
    g_ev_stop = nullptr;
  }
}
void note_launch() { ++g_launches; }
void reset_launch_count() { g_launches = 0; }

int launch_attn_simt(const HiAttnArgs& args, cudaStream_t stream);
int launch_attn_tc(const HiAttnArgs& args, cudaStream_t stream);
bool attn_tc_supported(const HiAttnArgs& args);
int launch_attn_decode_tc(const HiAttnArgs& args, cudaStream_t stream);
bool attn_decode_tc_supported(const HiAttnArgs& args);
int launch_attn_pair(const HiAttnArgs& args, cudaStream_t stream);
bool attn_pair_supported(const HiAttnArgs& args);
int launch_varlen_pair(const HiVarlenArgs& args, cudaStream_t stream);
int64_t simt_workspace_bytes(int head_dim);

// MHA / group-2 decode: the TMA-fed swapped-operand kernel also streams KV a little faster than the cp.async kernel once the launch
// is several waves of uniform rows - graph-timed on B200, 32q/32kv heads at ctx 2048: batch 64 0.3003 vs 0.3083 ms (7.15 TB/s =
// 109 % of the measured copy bandwidth), batch 128 0.595 vs 0.601, batch 64 at ctx 4096 0.593 vs 0.602, 32q/16kv batch 64 0.160 vs
// 0.167 - and loses where its one-CTA-per-(row, head) grid is a partial wave or the rows are ragged (batch 16 at ctx 2048 0.092 vs
// 0.085, batch 16 at ctx 8192 0.356 vs 0.308, 64 ragged rows 0.310 vs 0.304, ctx 512 0.086 vs 0.085): those stay on the cp.async kernel.
static bool decode_tc_wins_ungrouped(const HiAttnArgs& a) {
  const int64_t ctas = static_cast<int64_t>(a.n_tokens) * a.n_kv_heads;  // one CTA per (row, KV head), 3 resident per SM
  if (ctas < 4 * 3 * 148) return false;
  if (a.max_kv_len < 1024 || a.max_kv_len > 6144) return false;
  if (a.kv_blocks_hint <= 0 || a.n_seqs <= 0) return false;  // no length information: cannot tell uniform from ragged
  const double mean_len = static_cast<double>(a.kv_blocks_hint) * a.block_size / a.n_seqs;
  return a.max_kv_len <= 1.15 * mean_len;
}

}  // namespace hi

extern "C" const char* hi_last_error(void) { return hi::g_error; }
extern "C" int hi_abi_version(void) { return HI_B200_ABI_VERSION; }
extern "C" int hi_last_launch_count(void) { return hi::g_launches; }

extern "C" int64_t hi_attention_workspace_bytes(int32_t n_tokens, int32_t n_qo_heads, int32_t head_dim, int32_t max_kv_len) {
  // Upper bound over every kernel's split rule.  The finest any of them cuts is 256 keys per split (pair / tile kernels; the
  // decode kernels stop at 512) and the partials of one launch never exceed kMaxPartialBytes (common.cuh); on top come the
  // device-built plan (at most n_seqs + n_tokens / tile <= 2 * n_tokens entries of 8 bytes, the launcher wants room for it
  // twice plus 1 MiB) and the work counter in the tail.  Non-positive extents ask for the bound of ANY launch.
  using namespace hi;
  const int d = head_dim > 0 ? head_dim : 128;
  int64_t partial = kMaxPartialBytes;
  if (n_tokens > 0 && n_qo_heads > 0 && max_kv_len > 0) {
    const int64_t max_splits = (static_cast<int64_t>(max_kv_len) + 255) / 256;
    const int64_t per_split = partial_bytes_per_split(n_tokens, n_qo_heads, d);
    if (per_split > kMaxPartialBytes) partial = 0;                       // cap_splits() leaves such a launch unsplit
    else if (per_split * max_splits < kMaxPartialBytes) partial = per_split * max_splits;
  }
  if (partial < simt_workspace_bytes(d)) partial = simt_workspace_bytes(d);  // the CUDA-core kernel's own bound (attn_simt.cu)
  const int64_t plan = n_tokens > 0 ? 2 * ((static_cast<int64_t>(n_tokens) * 16 + 255) & ~int64_t(255)) : (int64_t(2) << 20);
  const int64_t need = partial + plan + (int64_t(1) << 20) + kWorkspaceTailBytes;
  return (need + 255) / 256 * 256;
}

extern "C" int32_t hi_attention_tile_tokens(int32_t n_qo_heads, int32_t n_kv_heads) {
  if (n_qo_heads <= 0 || n_kv_heads <= 0 || n_qo_heads % n_kv_heads != 0) return 0;
  const int group = n_qo_heads / n_kv_heads;
  return group <= 128 ? 2 * (128 / group) : 0;  // two 128-row tiles of (token, head-in-group) rows per CTA (attn_tc2.cu)
}

extern "C" int hi_paged_attention(const HiAttnArgs* p, void* stream_) {
  using namespace hi;
  reset_launch_count();
  HI_CHECK_ARG(p != nullptr, "paged_attention: null args");
  const HiAttnArgs& a = *p;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HI_CHECK_ARG(a.n_seqs >= 0 && a.n_tokens >= 0, "paged_attention: negative extents");
  if (a.n_seqs == 0 || a.n_tokens == 0) return HI_OK;
  HI_CHECK_ARG(a.q && a.out && a.key_cache && a.value_cache, "paged_attention: null tensor pointer");
  HI_CHECK_ARG(a.q_cu_seq_lens && a.kv_cu_seq_lens && a.block_tables && a.cu_blocks_lens,
               "paged_attention: null metadata pointer");
  HI_CHECK_ARG(a.n_qo_heads > 0 && a.n_kv_heads > 0 && a.n_qo_heads % a.n_kv_heads == 0,
               "paged_attention: n_qo_heads %d is not divisible by n_kv_heads %d", a.n_qo_heads, a.n_kv_heads);
  HI_CHECK_ARG(a.block_size > 0 && a.head_dim > 0, "paged_attention: bad block_size %d / head_dim %d", a.block_size, a.head_dim);
  HI_CHECK_ARG(a.max_q_len >= 1 && a.max_kv_len >= a.max_q_len, "paged_attention: bad max lens q=%d kv=%d", a.max_q_len, a.max_kv_len);
  const int es = dtype_size(a.dtype);
  HI_CHECK_SUPPORTED(es != 0, "paged_attention: unsupported dtype %d", a.dtype);
  const int64_t row = static_cast<int64_t>(a.n_qo_heads) * a.head_dim;
  HI_CHECK_ARG(a.q_row_stride >= row && a.out_row_stride >= row, "paged_attention: row stride smaller than n_qo_heads*head_dim");
  // Every lane-level access is at least 8 bytes wide.
  HI_CHECK_ARG(aligned_to(a.q, 16) && aligned_to(a.out, 16) && aligned_to(a.key_cache, 16) && aligned_to(a.value_cache, 16),
               "paged_attention: tensors must be 16-byte aligned");
  HI_CHECK_ARG((a.q_row_stride * es) % 16 == 0 && (a.out_row_stride * es) % 16 == 0,
               "paged_attention: row strides must be multiples of 16 bytes");
  HI_DEVICE_GUARD(a.device);

  int path = a.path;
  if (const char* env = tuning_env("HI_ATTN_PATH")) {  // test / profiling override: "simt" or "tc"
    if (env[0] == 's') path = HI_ATTN_SIMT;
    if (env[0] == 't') path = HI_ATTN_TCGEN05;
    if (env[0] == 'd') path = HI_ATTN_TCGEN05_DECODE;
    if (env[0] == 'p') path = HI_ATTN_TCGEN05_PAIR;
  }
  if (a.options != 0) {
    // mha_varlen_fwd's score options run on the any-shape CUDA-core kernel only (the paged attention layer never passes them)
    HI_CHECK_ARG((a.options & ~(HI_ATTN_OPT_WINDOW | HI_ATTN_OPT_SOFTCAP | HI_ATTN_OPT_ALIBI)) == 0, "paged_attention: unknown option bits 0x%x", a.options);
    HI_CHECK_ARG(!(a.options & HI_ATTN_OPT_SOFTCAP) || a.softcap > 0.f, "paged_attention: softcap must be positive, got %f", static_cast<double>(a.softcap));
    HI_CHECK_ARG(!(a.options & HI_ATTN_OPT_ALIBI) || a.alibi_slopes != nullptr, "paged_attention: HI_ATTN_OPT_ALIBI without alibi_slopes");
    HI_CHECK_SUPPORTED(path == HI_ATTN_AUTO || path == HI_ATTN_SIMT,
                       "paged_attention: softcap / sliding window / alibi run on the CUDA-core path only (path %d requested)", path);
    path = HI_ATTN_SIMT;
  }
  if (path == HI_ATTN_AUTO) {
    // Rows with q_len > 1 are dense contractions: tensor pipe.  Pure decode batches stream KV once per row: the
    // split-KV kernel keeps more bytes in flight and balances ragged lengths.
    // Grouped models (>= 4 query heads per KV head) decode on tensor cores too: the CUDA-core kernel needs G FMAs per KV
    // element and stops being memory-bound.  Decode-only batches take the swapped-operand kernel (tokens on the MMA M
    // side, one row per CTA); batches with prefill rows take the tile kernel for every row.
    const int group = a.n_qo_heads / a.n_kv_heads;
    if (a.max_q_len > 1) {
      path = attn_pair_supported(a) ? HI_ATTN_TCGEN05_PAIR : attn_tc_supported(a) ? HI_ATTN_TCGEN05 : HI_ATTN_SIMT;
    } else if (a.head_dim != 64 && a.head_dim != 128 && a.head_dim != 256 && attn_pair_supported(a)) {
      path = HI_ATTN_TCGEN05_PAIR;  // head dims only the tile kernel covers (96, ...): it takes decode rows as one-token tiles
    } else if (group >= 4 && attn_decode_tc_supported(a)) {
      path = HI_ATTN_TCGEN05_DECODE;
    } else if (decode_tc_wins_ungrouped(a) && attn_decode_tc_supported(a)) {
      path = HI_ATTN_TCGEN05_DECODE;
    } else if (group >= 4 && attn_tc_supported(a)) {
      path = HI_ATTN_TCGEN05;
    } else {
      path = HI_ATTN_SIMT;
    }
  }
  if (path == HI_ATTN_TCGEN05) return launch_attn_tc(a, stream);
  if (path == HI_ATTN_TCGEN05_DECODE) return launch_attn_decode_tc(a, stream);
  if (path == HI_ATTN_TCGEN05_PAIR) return launch_attn_pair(a, stream);
  if (path == HI_ATTN_SIMT) return launch_attn_simt(a, stream);
  set_error("paged_attention: unknown path %d", path);
  return HI_ERR_INVALID_ARGUMENT;
}

extern "C" int hi_varlen_attention(const HiVarlenArgs* p, void* stream_) {
  using namespace hi;
  reset_launch_count();
  HI_CHECK_ARG(p != nullptr, "varlen_attention: null args");
  const HiVarlenArgs& a = *p;
  HI_CHECK_ARG(a.n_seqs >= 0 && a.n_q_tokens >= 0 && a.n_k_tokens >= 0, "varlen_attention: negative extents");
  if (a.n_seqs == 0 || a.n_q_tokens == 0) return HI_OK;
  HI_CHECK_ARG(a.q && a.k && a.v && a.out && a.cu_seqlens_q && a.cu_seqlens_k, "varlen_attention: null pointer");
  HI_CHECK_ARG(a.n_qo_heads > 0 && a.n_kv_heads > 0 && a.n_qo_heads % a.n_kv_heads == 0,
               "varlen_attention: n_qo_heads %d is not divisible by n_kv_heads %d", a.n_qo_heads, a.n_kv_heads);
  HI_CHECK_ARG(a.head_dim > 0 && a.max_q_len >= 1 && a.max_kv_len >= 1, "varlen_attention: bad head_dim %d / max lens q=%d kv=%d",
               a.head_dim, a.max_q_len, a.max_kv_len);
  HI_CHECK_SUPPORTED(dtype_size(a.dtype) != 0, "varlen_attention: unsupported dtype %d", a.dtype);
  const int64_t qrow = static_cast<int64_t>(a.n_qo_heads) * a.head_dim, krow = static_cast<int64_t>(a.n_kv_heads) * a.head_dim;
  HI_CHECK_ARG(a.q_row_stride >= qrow && a.out_row_stride >= qrow && a.k_row_stride >= krow && a.v_row_stride >= krow,
               "varlen_attention: row stride smaller than n_heads*head_dim");
  HI_DEVICE_GUARD(a.device);
  return launch_varlen_pair(a, static_cast<cudaStream_t>(stream_));
}

extern "C" int hi_event_create(void** event_out) {
  using namespace hi;
  HI_CHECK_ARG(event_out != nullptr, "event_create: null output");
  cudaEvent_t ev;
  HI_CUDA(cudaEventCreate(&ev));
  *event_out = ev;
  return HI_OK;
}
extern "C" int hi_event_destroy(void* event) {
  using namespace hi;
  if (event != nullptr) HI_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return HI_OK;
}
extern "C" int hi_event_record(void* event, void* stream) {
  using namespace hi;
  HI_CHECK_ARG(event != nullptr, "event_record: null event");
  HI_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)));
  return HI_OK;
}
extern "C" int hi_event_elapsed_ms(void* start, void* stop, float* ms_out) {
  using namespace hi;
  HI_CHECK_ARG(start && stop && ms_out, "event_elapsed_ms: null argument");
  HI_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(stop)));
  HI_CUDA(cudaEventElapsedTime(ms_out, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  return HI_OK;
}
extern "C" int hi_set_kernel_timing_events(void* start, void* stop) {
  hi::g_ev_start = static_cast<cudaEvent_t>(start);
  hi::g_ev_stop = static_cast<cudaEvent_t>(stop);
  return HI_OK;
}
