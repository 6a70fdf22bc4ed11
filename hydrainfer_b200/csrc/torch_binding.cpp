// Compiled Python boundary: the reference's five pybind modules, same names and positional signatures, over the C ABI of
// include/hi_b200.h.
//
//   PyInit_kv_cache_kernels     set_kv_cache                       csrc/kernel/kv_cache_kernels/kv_cache_kernels_pybind.cpp:7-10
//   PyInit_cache_kernels        set_image_cache (+ get_image_cache) csrc/kernel/cache_kernels/cache_kernels_pybind.cpp:7-10
//   PyInit_position_embedding   apply_rotary_pos_emb (+ rope_set_kv_cache)  csrc/kernel/position_embedding/position_embedding_pybind.cpp
//   PyInit_flash_attn           mha_varlen_fwd (+ append_and_attend)        csrc/kernel/flash_attn/flash_attn_pybind.cpp, flash_api.cpp:216-355,
//                                                                   stub hydrainfer/_C/kernel/flash_attn/__init__.pyi:22-40
//   PyInit_block_migration      get_ipc_mem_handle, register_ipc_mem_handle, migrate_blocks (+ migrate_blocks_layers, push_blocks)
//                                                                   csrc/data_transfer/block_migration_pybind.cpp:10-15
//
// One translation unit defines all five module inits; hydrainfer_b200/build.py links it once and installs the shared object
// under each module's file name (hydrainfer_b200/_C/kernel/<name><EXT_SUFFIX>, _C/data_transfer/<name><EXT_SUFFIX>), where the
// reference's CMake puts its own (csrc/CMakeLists.txt:4-11).  No kernel lives here: every function checks its arguments the
// way the reference's CHECK / TORCH_CHECK lines do (raising RuntimeError instead of aborting), takes torch's current stream of
// the tensors' device (at::cuda::getCurrentCUDAStream() in the reference) and calls libhi_b200.so.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAStream.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <torch/csrc/utils/pybind.h>

#include <cstring>
#include <map>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hi_b200.h"

namespace py = pybind11;
using at::Tensor;

namespace {

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error(msg); }

void check(int status) {
  if (status != HI_OK) fail("hi_b200 error " + std::to_string(status) + ": " + hi_last_error());
}

int dtype_code(const Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return HI_F32;
    case at::kHalf: return HI_F16;
    case at::kBFloat16: return HI_BF16;
    default: fail(std::string("hi_b200: dtype ") + c10::toString(t.scalar_type()) + " is not supported (float32, float16, bfloat16)");
  }
}

// All tensors on one CUDA device; returns its index.  There is no CPU fallback in this package.
int require_cuda(std::initializer_list<const Tensor*> tensors) {
  int dev = -1;
  for (const Tensor* t : tensors) {
    if (!t->defined()) continue;
    if (!t->is_cuda()) fail("hi_b200: the CUDA extension only accepts CUDA tensors; there is no CPU fallback");
    const int d = static_cast<int>(t->get_device());
    if (dev < 0) dev = d;
    if (d != dev) fail("hi_b200: tensors on different devices (cuda:" + std::to_string(d) + " vs cuda:" + std::to_string(dev) + ")");
  }
  if (dev < 0) fail("hi_b200: no tensor argument");
  return dev;
}

void* current_stream(int device) { return c10::cuda::getCurrentCUDAStream(static_cast<c10::DeviceIndex>(device)).stream(); }

bool heads_contiguous(const Tensor& t) { return t.stride(-1) == 1 && t.stride(-2) == t.size(-1); }
int64_t row_stride(const Tensor& t) { return t.size(0) > 1 ? t.stride(0) : t.size(1) * t.size(2); }

void check_slots(const char* fn, const Tensor& slot_ids, int64_t n_tokens) {
  if (slot_ids.scalar_type() != at::kInt || slot_ids.dim() != 1 || !slot_ids.is_contiguous() || slot_ids.size(0) != n_tokens)
    fail(std::string(fn) + ": slot_ids must be a contiguous int32 vector with one entry per token");
}

// ---- kv_cache_kernels ----------------------------------------------------------------------------------------------------------
// key_cache[slot / bs, slot % bs] = keys[t]; same for values (kv_cache_kernels.cu:60-95).  Layout violations raise (the
// reference CHECK-aborts, :67-68).
void set_kv_cache(const Tensor& slot_ids, const Tensor& keys, const Tensor& values, Tensor& key_cache, Tensor& value_cache) {
  const int dev = require_cuda({&slot_ids, &keys, &values, &key_cache, &value_cache});
  if (keys.dim() != 3 || values.sizes() != keys.sizes())
    fail("set_kv_cache: keys/values must be [n_tokens, n_heads, head_dim] of equal shape");
  if (!heads_contiguous(keys) || !heads_contiguous(values)) fail("set_kv_cache: keys and values must be contiguous over (n_heads, head_dim)");
  if (key_cache.dim() != 4 || key_cache.sizes() != value_cache.sizes() || !key_cache.is_contiguous() || !value_cache.is_contiguous())
    fail("set_kv_cache: caches must be contiguous [n_blocks, block_size, n_heads, head_dim] of equal shape");
  if (key_cache.size(2) != keys.size(1) || key_cache.size(3) != keys.size(2)) fail("set_kv_cache: cache heads/dim differ from keys");
  if (keys.scalar_type() != values.scalar_type() || keys.scalar_type() != key_cache.scalar_type() || keys.scalar_type() != value_cache.scalar_type())
    fail("set_kv_cache: dtype mismatch between keys, values and caches");
  check_slots("set_kv_cache", slot_ids, keys.size(0));
  check(hi_set_kv_cache(slot_ids.data_ptr<int32_t>(), keys.data_ptr(), values.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(),
                        keys.size(0), keys.size(1) * keys.size(2), row_stride(keys), row_stride(values), dtype_code(keys), dev,
                        current_stream(dev)));
}

// ---- cache_kernels -------------------------------------------------------------------------------------------------------------
void set_image_cache(const Tensor& slot_ids, const Tensor& image_tokens, Tensor& image_cache) {
  const int dev = require_cuda({&slot_ids, &image_tokens, &image_cache});
  if (image_tokens.dim() != 3 || !heads_contiguous(image_tokens))
    fail("set_image_cache: image_tokens must be [n_tokens, n_heads, head_dim], contiguous over the last two dims");
  if (image_cache.dim() != 4 || !image_cache.is_contiguous() || image_cache.size(2) != image_tokens.size(1) || image_cache.size(3) != image_tokens.size(2))
    fail("set_image_cache: image_cache must be contiguous [n_blocks, block_size, n_heads, head_dim] matching the tokens");
  if (image_tokens.scalar_type() != image_cache.scalar_type()) fail("set_image_cache: dtype mismatch");
  check_slots("set_image_cache", slot_ids, image_tokens.size(0));
  check(hi_set_image_cache(slot_ids.data_ptr<int32_t>(), image_tokens.data_ptr(), image_cache.data_ptr(), image_tokens.size(0),
                           image_tokens.size(1) * image_tokens.size(2), row_stride(image_tokens), dtype_code(image_tokens), dev,
                           current_stream(dev)));
}

// image_cache.view(-1, H * d)[slot_ids, :] -> [T, H * d]: the read side of the image-embedding cache
// (hydrainfer/engine/parameters_builder.py:48-55).  Extension of the reference module.
Tensor get_image_cache(const Tensor& slot_ids, const Tensor& image_cache) {
  const int dev = require_cuda({&slot_ids, &image_cache});
  if (image_cache.dim() != 4 || !image_cache.is_contiguous()) fail("get_image_cache: image_cache must be contiguous [n_blocks, block_size, n_heads, head_dim]");
  if (slot_ids.scalar_type() != at::kInt || slot_ids.dim() != 1 || !slot_ids.is_contiguous()) fail("get_image_cache: slot_ids must be a contiguous int32 vector");
  const int64_t n = slot_ids.size(0), row = image_cache.size(2) * image_cache.size(3);
  Tensor out = at::empty({n, row}, image_cache.options());
  check(hi_get_image_cache(slot_ids.data_ptr<int32_t>(), image_cache.data_ptr(), out.data_ptr(), n, row, row, dtype_code(image_cache), dev,
                           current_stream(dev)));
  return out;
}

// ---- position_embedding --------------------------------------------------------------------------------------------------------
void check_rope(const char* fn, const Tensor& q, const Tensor& k, const Tensor& positions, const Tensor& cos_sin, int64_t rotary_dim) {
  const std::string name(fn);
  if (q.dim() != 3 || k.dim() != 3 || q.size(0) != k.size(0) || q.size(2) != k.size(2))
    fail(name + ": query/key must be [n_tokens, n_heads, head_dim] with equal n_tokens and head_dim");
  if (!heads_contiguous(q) || !heads_contiguous(k)) fail(name + ": query and key must be contiguous over (n_heads, head_dim)");  // rope.cu:100-101
  if (k.scalar_type() != q.scalar_type()) fail(name + ": query/key dtype mismatch");
  if (positions.dim() != 1 || positions.size(0) != q.size(0) || !positions.is_contiguous() ||
      (positions.scalar_type() != at::kInt && positions.scalar_type() != at::kLong))
    fail(name + ": positions must be a contiguous int32/int64 vector with one entry per token");
  if (rotary_dim % 2 != 0 || rotary_dim > q.size(2)) fail(name + ": rotary_dim must be even and <= head_dim");
  if (!cos_sin.is_contiguous() || cos_sin.dim() != 3 || cos_sin.size(1) != 2 || cos_sin.size(2) * 2 != rotary_dim)
    fail(name + ": cos_sin must be contiguous [max_positions, 2, rotary_dim/2]");
  if (cos_sin.scalar_type() != q.scalar_type() && cos_sin.scalar_type() != at::kFloat)
    fail(name + ": cos_sin must be float32 or have the query dtype");
}

void launch_rope(int dev, Tensor& q, Tensor& k, const Tensor* v, const Tensor& positions, const Tensor& cos_sin, int64_t rotary_dim, bool interleaved,
                 const Tensor* slot_ids, Tensor* key_cache, Tensor* value_cache, bool write_back_k, bool force_scalar) {
  HiRopeArgs a;
  std::memset(&a, 0, sizeof(a));
  a.q = q.data_ptr();
  a.k = k.data_ptr();
  a.q_row_stride = row_stride(q);
  a.k_row_stride = row_stride(k);
  a.positions = positions.data_ptr();
  a.cos_sin = cos_sin.data_ptr();
  if (slot_ids != nullptr) {
    a.v = v->data_ptr();
    a.v_row_stride = row_stride(*v);
    a.slot_ids = slot_ids->data_ptr<int32_t>();
    a.key_cache = key_cache->data_ptr();
    a.value_cache = value_cache->data_ptr();
  }
  a.n_tokens = q.size(0);
  a.n_qo_heads = static_cast<int32_t>(q.size(1));
  a.n_kv_heads = static_cast<int32_t>(k.size(1));
  a.head_dim = static_cast<int32_t>(q.size(2));
  a.rotary_dim = static_cast<int32_t>(rotary_dim);
  a.dtype = dtype_code(q);
  a.cos_sin_dtype = dtype_code(cos_sin);
  a.positions_int64 = positions.scalar_type() == at::kLong ? 1 : 0;
  a.interleaved = interleaved ? 1 : 0;
  a.write_back_k = write_back_k ? 1 : 0;
  a.force_scalar = force_scalar ? 1 : 0;
  a.device = dev;
  check(hi_rope_append(&a, current_stream(dev)));
}

// Rotates query [T, Hq, d] and key [T, Hkv, d] in place (rope.cu:90-117).
void apply_rotary_pos_emb(Tensor& query, Tensor& key, const Tensor& positions, const Tensor& cos_sin, int64_t rotary_dim, bool interleaved) {
  const int dev = require_cuda({&query, &key, &positions, &cos_sin});
  check_rope("apply_rotary_pos_emb", query, key, positions, cos_sin, rotary_dim);
  launch_rope(dev, query, key, nullptr, positions, cos_sin, rotary_dim, interleaved, nullptr, nullptr, nullptr, true, false);
}

// apply_rotary_pos_emb + set_kv_cache in one launch (extension): query rotated in place, rotated key and the value go to their
// cache slots; key itself is rewritten only if write_back_k.
void rope_set_kv_cache(Tensor& query, Tensor& key, const Tensor& value, const Tensor& positions, const Tensor& cos_sin, int64_t rotary_dim,
                       bool interleaved, const Tensor& slot_ids, Tensor& key_cache, Tensor& value_cache, bool write_back_k, bool force_scalar) {
  const int dev = require_cuda({&query, &key, &value, &positions, &cos_sin, &slot_ids, &key_cache, &value_cache});
  check_rope("rope_set_kv_cache", query, key, positions, cos_sin, rotary_dim);
  if (value.sizes() != key.sizes() || value.scalar_type() != key.scalar_type() || !heads_contiguous(value))
    fail("rope_set_kv_cache: value must have the key's shape and dtype and be contiguous over (n_heads, head_dim)");
  if (key_cache.dim() != 4 || key_cache.sizes() != value_cache.sizes() || !key_cache.is_contiguous() || !value_cache.is_contiguous())
    fail("rope_set_kv_cache: caches must be contiguous [n_blocks, block_size, n_heads, head_dim] of equal shape");
  if (key_cache.size(2) != key.size(1) || key_cache.size(3) != key.size(2) || key_cache.scalar_type() != key.scalar_type() ||
      value_cache.scalar_type() != key.scalar_type())
    fail("rope_set_kv_cache: cache geometry / dtype differs from the keys");
  check_slots("rope_set_kv_cache", slot_ids, key.size(0));
  launch_rope(dev, query, key, &value, positions, cos_sin, rotary_dim, interleaved, &slot_ids, &key_cache, &value_cache, write_back_k, force_scalar);
}

// ---- flash_attn ----------------------------------------------------------------------------------------------------------------
// Split-KV scratch: persistent (like flashinfer's workspace, hydrainfer/engine/executor.py:99), one buffer per (device, stream)
// so two attention calls in flight on different streams of one device never share partials or the work counter; grown to
// hi_attention_workspace_bytes() of the largest launch seen.  The reference allocates its scratch per call (flash_api.cpp:311,
// 330-337).
std::mutex g_ws_mu;
std::map<std::pair<int, void*>, Tensor> g_workspaces;

Tensor& workspace_for(int dev, void* stream, int64_t need) {
  std::lock_guard<std::mutex> lock(g_ws_mu);
  Tensor& ws = g_workspaces[{dev, stream}];
  if (!ws.defined() || ws.numel() < need) {
    int64_t size = need;
    if (ws.defined() && 2 * ws.numel() > size) size = 2 * ws.numel();
    ws = at::empty({size}, at::TensorOptions().dtype(at::kByte).device(at::kCUDA, static_cast<c10::DeviceIndex>(dev)));
  }
  return ws;
}

void check_i32(const char* what, const Tensor& t) {
  if (t.scalar_type() != at::kInt || !t.is_contiguous()) fail(std::string("mha_varlen_fwd: ") + what + " must be a contiguous int32 tensor");
}

void varlen_fwd(Tensor& out, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& cu_q, const Tensor& cu_k, int64_t max_q, int64_t max_k,
                double scale, bool causal) {
  const int dev = require_cuda({&out, &q, &k, &v, &cu_q, &cu_k});
  for (const Tensor* t : {&q, &k, &v, static_cast<const Tensor*>(&out)})
    if (t->dim() != 3 || !heads_contiguous(*t)) fail("mha_varlen_fwd: q, k, v and out must be [n_tokens, n_heads, head_dim], contiguous over the last two dims");
  if (out.sizes() != q.sizes() || k.sizes() != v.sizes() || k.size(2) != q.size(2) || q.size(1) % k.size(1) != 0) fail("mha_varlen_fwd: shape mismatch between q, out, k and v");
  if (q.scalar_type() != out.scalar_type() || q.scalar_type() != k.scalar_type() || q.scalar_type() != v.scalar_type()) fail("mha_varlen_fwd: dtype mismatch");
  if (q.scalar_type() != at::kHalf && q.scalar_type() != at::kBFloat16) fail("mha_varlen_fwd: only fp16 and bf16 are supported");  // flash_api.cpp:236
  check_i32("cu_seqlens_q", cu_q);
  check_i32("cu_seqlens_k", cu_k);
  const int64_t n_seqs = cu_q.size(0) - 1;
  if (cu_q.dim() != 1 || cu_k.dim() != 1 || cu_k.size(0) != n_seqs + 1) fail("mha_varlen_fwd: cu_seqlens_q and cu_seqlens_k must both have batch + 1 entries");
  void* stream = current_stream(dev);
  Tensor& ws = workspace_for(dev, stream, 4096);
  HiVarlenArgs a;
  std::memset(&a, 0, sizeof(a));
  a.q = q.data_ptr();
  a.k = k.data_ptr();
  a.v = v.data_ptr();
  a.out = out.data_ptr();
  a.q_row_stride = row_stride(q);
  a.k_row_stride = row_stride(k);
  a.v_row_stride = row_stride(v);
  a.out_row_stride = row_stride(out);
  a.cu_seqlens_q = cu_q.data_ptr<int32_t>();
  a.cu_seqlens_k = cu_k.data_ptr<int32_t>();
  a.n_seqs = static_cast<int32_t>(n_seqs);
  a.n_q_tokens = static_cast<int32_t>(q.size(0));
  a.n_k_tokens = static_cast<int32_t>(k.size(0));
  a.max_q_len = static_cast<int32_t>(max_q);
  a.max_kv_len = static_cast<int32_t>(max_k);
  a.n_qo_heads = static_cast<int32_t>(q.size(1));
  a.n_kv_heads = static_cast<int32_t>(k.size(1));
  a.head_dim = static_cast<int32_t>(q.size(2));
  a.dtype = dtype_code(q);
  a.causal = causal ? 1 : 0;
  a.softmax_scale = static_cast<float>(scale);
  a.device = dev;
  a.workspace = ws.data_ptr();
  a.workspace_bytes = ws.numel();
  check(hi_varlen_attention(&a, stream));
}

// mha_varlen_fwd's score options (flash_api.cpp:225-232); the default is what the paged attention layer passes: causal, nothing else
struct ScoreOptions {
  const Tensor* alibi_slopes = nullptr;
  double softcap = 0.0;
  int64_t window_left = -1, window_right = 0;
};

void paged_fwd(int dev, Tensor& out, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& cu_q, const Tensor& cu_k, const Tensor& block_table,
               const Tensor& cu_block_lens, int64_t max_q, int64_t max_k, double scale, int64_t path, const std::optional<Tensor>& work_items,
               int64_t work_tile_tokens, int64_t qk_work_hint, const ScoreOptions& opt = ScoreOptions()) {
  if (q.dim() != 3 || out.sizes() != q.sizes() || !heads_contiguous(q) || !heads_contiguous(out))
    fail("mha_varlen_fwd: q and out must be [n_tokens, n_heads, head_dim], contiguous over the last two dims");
  if (k.dim() != 4 || k.sizes() != v.sizes() || !k.is_contiguous() || !v.is_contiguous())
    fail("mha_varlen_fwd: k and v must be contiguous paged caches [n_blocks, block_size, n_kv_heads, head_dim]");
  if (out.scalar_type() != q.scalar_type() || k.scalar_type() != q.scalar_type() || v.scalar_type() != q.scalar_type()) fail("mha_varlen_fwd: dtype mismatch");
  check_i32("cu_seqlens_q", cu_q);
  check_i32("cu_seqlens_k", cu_k);
  check_i32("block_table", block_table);
  check_i32("cu_block_lens", cu_block_lens);
  const int64_t n_tokens = q.size(0), n_qo_heads = q.size(1), head_dim = q.size(2);
  const int64_t n_blocks = k.size(0), block_size = k.size(1), n_kv_heads = k.size(2);
  if (k.size(3) != head_dim || n_kv_heads == 0 || n_qo_heads % n_kv_heads != 0) fail("mha_varlen_fwd: head mismatch between q and the caches");
  const int64_t n_seqs = cu_q.size(0) - 1;
  if (cu_k.size(0) != n_seqs + 1 || cu_block_lens.size(0) != n_seqs + 1) fail("mha_varlen_fwd: cu_seqlens_q, cu_seqlens_k and cu_block_lens must all have batch + 1 entries");
  if (work_items.has_value() && (work_items->scalar_type() != at::kInt || !work_items->is_cuda() || work_items->get_device() != dev || work_items->dim() != 2 ||
                                 work_items->size(1) != 2 || !work_items->is_contiguous()))
    fail("mha_varlen_fwd: work_items must be a contiguous int32 device tensor of shape [n_items, 2]");
  void* stream = current_stream(dev);
  Tensor& ws = workspace_for(dev, stream, hi_attention_workspace_bytes(static_cast<int32_t>(n_tokens), static_cast<int32_t>(n_qo_heads),
                                                                       static_cast<int32_t>(head_dim), static_cast<int32_t>(max_k)));
  HiAttnArgs a;
  std::memset(&a, 0, sizeof(a));
  a.q = q.data_ptr();
  a.out = out.data_ptr();
  a.key_cache = k.data_ptr();
  a.value_cache = v.data_ptr();
  a.q_row_stride = n_tokens > 1 ? q.stride(0) : n_qo_heads * head_dim;
  a.out_row_stride = n_tokens > 1 ? out.stride(0) : n_qo_heads * head_dim;
  a.q_cu_seq_lens = cu_q.data_ptr<int32_t>();
  a.kv_cu_seq_lens = cu_k.data_ptr<int32_t>();
  a.block_tables = block_table.data_ptr<int32_t>();
  a.cu_blocks_lens = cu_block_lens.data_ptr<int32_t>();
  a.n_seqs = static_cast<int32_t>(n_seqs);
  a.n_tokens = static_cast<int32_t>(n_tokens);
  a.max_q_len = static_cast<int32_t>(max_q);
  a.max_kv_len = static_cast<int32_t>(max_k);
  a.n_qo_heads = static_cast<int32_t>(n_qo_heads);
  a.n_kv_heads = static_cast<int32_t>(n_kv_heads);
  a.head_dim = static_cast<int32_t>(head_dim);
  a.block_size = static_cast<int32_t>(block_size);
  a.n_blocks = n_blocks;
  a.dtype = dtype_code(q);
  a.softmax_scale = static_cast<float>(scale);
  a.workspace = ws.data_ptr();
  a.workspace_bytes = ws.numel();
  a.path = static_cast<int32_t>(path);
  a.device = dev;
  a.kv_blocks_hint = static_cast<int32_t>(block_table.numel());
  if (work_items.has_value()) {
    a.work_items = work_items->data_ptr<int32_t>();
    a.n_work_items = static_cast<int32_t>(work_items->size(0));
  }
  a.qk_work_hint = qk_work_hint;
  a.work_tile_tokens = static_cast<int32_t>(work_tile_tokens);
  if (!(opt.window_left < 0 && opt.window_right == 0)) {  // anything but plain causal (flash_api.cpp:104-111)
    a.options |= HI_ATTN_OPT_WINDOW;
    a.window_left = static_cast<int32_t>(opt.window_left < 0 ? -1 : opt.window_left);
    a.window_right = static_cast<int32_t>(opt.window_right < 0 ? -1 : opt.window_right);
  }
  if (opt.softcap > 0.0) {
    a.options |= HI_ATTN_OPT_SOFTCAP;
    a.softcap = static_cast<float>(opt.softcap);
  }
  if (opt.alibi_slopes != nullptr) {  // flash_api.cpp:202-209
    const Tensor& al = *opt.alibi_slopes;
    if (al.scalar_type() != at::kFloat) fail("mha_varlen_fwd: ALiBi slopes must have dtype fp32");
    if (!al.is_cuda() || al.get_device() != dev) fail("mha_varlen_fwd: alibi_slopes must be on the device of q");
    if (al.stride(-1) != 1) fail("mha_varlen_fwd: ALiBi slopes tensor must have contiguous last dimension");
    const bool per_head = al.dim() == 1 && al.size(0) == n_qo_heads;
    const bool per_seq = al.dim() == 2 && al.size(0) == n_seqs && al.size(1) == n_qo_heads;
    if (!per_head && !per_seq) fail("mha_varlen_fwd: alibi_slopes must be [num_heads] or [batch_size, num_heads]");
    a.options |= HI_ATTN_OPT_ALIBI;
    a.alibi_slopes = al.data_ptr<float>();
    a.alibi_batch_stride = per_seq ? al.stride(0) : 0;
  }
  check(hi_paged_attention(&a, stream));
}

// The reference's 16 positional arguments (flash_api.cpp:216-232), writing `out` in place; `path`, `work_items`,
// `work_tile_tokens` and `qk_work_hint` are optional extensions after them (the host plan of AttentionParametersBuilder).
// Paged form (block_table + cu_block_lens, window (-1, 0) == causal: causal_attention.py:274-291) and un-paged form
// (block_table None, k / v [T, Hkv, d], window (-1, -1) or (-1, 0): multihead_attention.py:140-157, 194-211).  alibi, softcap and
// sliding windows are taken by the paged form (any-shape kernel); the un-paged form raises for them like TORCH_CHECK.
void mha_varlen_fwd(Tensor& out, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& cu_seqlens_q, const Tensor& cu_seqlens_k,
                    const std::optional<Tensor>& block_table, const std::optional<Tensor>& cu_block_lens, const std::optional<Tensor>& alibi_slopes,
                    int64_t max_seqlen_q, int64_t max_seqlen_k, double softmax_scale, double softcap, int64_t window_size_left, int64_t window_size_right,
                    int64_t num_splits, int64_t path, const std::optional<Tensor>& work_items, int64_t work_tile_tokens, int64_t qk_work_hint) {
  (void)num_splits;  // the split count is the kernels' own decision (as with num_splits == 0 in the reference, flash_api.cpp:330)
  if (!block_table.has_value()) {
    if (alibi_slopes.has_value() || softcap != 0 || window_size_left != -1 || (window_size_right != -1 && window_size_right != 0))
      fail("mha_varlen_fwd: alibi / softcap / sliding window are not implemented");
    varlen_fwd(out, q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, softmax_scale, window_size_right == 0);
    return;
  }
  if (!cu_block_lens.has_value()) fail("mha_varlen_fwd: a block_table needs cu_block_lens (flattened CSR block table)");
  // alibi / softcap / local windows: never passed by the paged attention layer (causal_attention.py:274-291), part of the entry point all
  // the same (the reference's FlashAttention-2 build has them enabled); they run on the any-shape CUDA-core kernel
  ScoreOptions opt;
  if (alibi_slopes.has_value()) opt.alibi_slopes = &*alibi_slopes;
  if (softcap < 0) fail("mha_varlen_fwd: softcap must be non-negative");
  opt.softcap = softcap;
  opt.window_left = window_size_left;
  opt.window_right = window_size_right;
  const int dev = require_cuda({&out, &q, &k, &v, &cu_seqlens_q, &cu_seqlens_k, &*block_table, &*cu_block_lens});
  paged_fwd(dev, out, q, k, v, cu_seqlens_q, cu_seqlens_k, *block_table, *cu_block_lens, max_seqlen_q, max_seqlen_k, softmax_scale, path, work_items,
            work_tile_tokens, qk_work_hint, opt);
}

// CausalGroupedQueryPageAttention.forward in one call (causal_attention.py:394-406): append the new K / V rows to the paged cache,
// then attend.  query [T, Hq, d], key / value [T, Hkv, d] (row strides free); returns o [T, Hq, d].  Extension: saves the
// second Python -> C++ transition of an eager layer call.
Tensor append_and_attend(const Tensor& query, const Tensor& key, const Tensor& value, const Tensor& new_cache_slots, Tensor& key_cache, Tensor& value_cache,
                         const Tensor& cu_seqlens_q, const Tensor& cu_seqlens_k, const Tensor& block_table, const Tensor& cu_block_lens, int64_t max_seqlen_q,
                         int64_t max_seqlen_k, double softmax_scale, int64_t path, const std::optional<Tensor>& work_items, int64_t work_tile_tokens,
                         int64_t qk_work_hint) {
  // query / key / value may also be the layer's 2-D [T, H * d] tensors (row strides free): the head geometry comes from the caches
  if (key_cache.dim() != 4) fail("append_and_attend: caches must be [n_blocks, block_size, n_kv_heads, head_dim]");
  const int64_t hkv = key_cache.size(2), d = key_cache.size(3);
  const bool flat = query.dim() == 2;
  if (flat && (d == 0 || query.size(1) % d != 0 || key.dim() != 2 || value.dim() != 2 || key.size(1) != hkv * d || value.size(1) != hkv * d))
    fail("append_and_attend: 2-D query / key / value must be [n_tokens, n_heads * head_dim] matching the caches");
  if (!flat && query.dim() != 3) fail("append_and_attend: query must be [n_tokens, n_heads, head_dim] or [n_tokens, n_heads * head_dim]");
  const int64_t n_tokens = query.size(0), hq = flat ? query.size(1) / d : query.size(1);
  const Tensor q3 = flat ? query.view({n_tokens, hq, d}) : query;
  const Tensor k3 = flat ? key.view({n_tokens, hkv, d}) : key;
  const Tensor v3 = flat ? value.view({n_tokens, hkv, d}) : value;
  set_kv_cache(new_cache_slots, k3, v3, key_cache, value_cache);
  const int dev = require_cuda({&q3, &key_cache, &cu_seqlens_q, &cu_seqlens_k, &block_table, &cu_block_lens});
  Tensor out = at::empty({n_tokens, hq, q3.size(2)}, query.options());
  paged_fwd(dev, out, q3, key_cache, value_cache, cu_seqlens_q, cu_seqlens_k, block_table, cu_block_lens, max_seqlen_q, max_seqlen_k, softmax_scale, path,
            work_items, work_tile_tokens, qk_work_hint);
  return flat ? out.view({n_tokens, hq * d}) : out;
}

int last_launch_count() { return hi_last_launch_count(); }

// The split-KV scratch of (device, torch's current stream of that device), at least `min_bytes` long: tests poison it, a CUDA-graph
// runner can size it for its largest captured batch before capturing.
Tensor workspace(int64_t device, int64_t min_bytes) {
  const int dev = static_cast<int>(device);
  return workspace_for(dev, current_stream(dev), min_bytes > 4096 ? min_bytes : hi_attention_workspace_bytes(0, 0, 128, 0));
}

// ---- block_migration -----------------------------------------------------------------------------------------------------------
// Differences from the reference module, all inside its contract: the peer pool is mapped once per process and cached (the
// reference re-opens the handle on every call, block_migration.cpp:213-215); all (layer, K/V, block) runs of a request move in
// one gather launch on the current stream (:222-244 issues n_layers * n_tokens * n_blocks cudaMemcpyAsync); a handle exported
// by THIS process resolves to the local pointer (cudaIpcOpenMemHandle cannot open a handle in the exporting process); a pool
// that does not start at the base of its allocation carries 8 extra ints with the byte offset (the reference assumes 0);
// errors raise RuntimeError instead of printf + exit(-1) (:9-15).
struct LocalExport {
  void* ptr;
  int device;
};
std::mutex g_bm_mu;
std::map<std::vector<int64_t>, LocalExport> g_local_exports;
std::vector<void*> g_registered;  // register_ipc_mem_handle returns an index into it (block_migration.cpp:61-80)

std::vector<int64_t> get_ipc_mem_handle(const Tensor& tensor) {
  const int dev = require_cuda({&tensor});
  uint8_t raw[64];
  int64_t offset = 0;
  check(hi_ipc_get_handle(tensor.data_ptr(), raw, &offset, dev));
  std::vector<int64_t> handle(raw, raw + 64);
  if (offset != 0)
    for (int i = 0; i < 8; ++i) handle.push_back((offset >> (8 * i)) & 0xff);
  std::lock_guard<std::mutex> lock(g_bm_mu);
  g_local_exports[handle] = LocalExport{tensor.data_ptr(), dev};
  return handle;
}

// Device pointer of the pool named by `handle`, usable from `device`; status of the C ABI on failure.
int resolve(const std::vector<int64_t>& handle, int device, void** out) {
  {
    std::lock_guard<std::mutex> lock(g_bm_mu);
    auto it = g_local_exports.find(handle);
    if (it != g_local_exports.end()) {
      if (it->second.device != device) {
        const int rc = hi_enable_peer_access(device, it->second.device);
        if (rc != HI_OK) return rc;
      }
      *out = it->second.ptr;
      return HI_OK;
    }
  }
  if (handle.size() != 64 && handle.size() != 72) fail("ipc handle must have 64 (or 72) entries, got " + std::to_string(handle.size()));
  uint8_t raw[64];
  for (int i = 0; i < 64; ++i) raw[i] = static_cast<uint8_t>(handle[i] & 0xff);
  int64_t offset = 0;
  if (handle.size() == 72)
    for (int i = 0; i < 8; ++i) offset |= (handle[64 + i] & 0xff) << (8 * i);
  return hi_ipc_open_handle(raw, offset, device, out);
}

int64_t register_ipc_mem_handle(const std::vector<int64_t>& kv_cache_handle_vec) {
  const int device = static_cast<int>(c10::cuda::current_device());
  void* ptr = nullptr;
  const int rc = resolve(kv_cache_handle_vec, device, &ptr);
  if (rc == HI_ERR_PEER_UNSUPPORTED) return -1;  // block_migration.cpp:73-76
  check(rc);
  std::lock_guard<std::mutex> lock(g_bm_mu);
  g_registered.push_back(ptr);
  return static_cast<int64_t>(g_registered.size()) - 1;
}

int64_t check_tables(const char* fn, const std::vector<int64_t>& src_bt, const std::vector<int64_t>& dst_bt, int64_t src_n_blocks, int64_t dst_n_blocks) {
  if (src_bt.size() != dst_bt.size())
    fail(std::string(fn) + ": block tables differ in length (" + std::to_string(src_bt.size()) + " vs " + std::to_string(dst_bt.size()) + ")");
  for (size_t i = 0; i < src_bt.size(); ++i)
    if (src_bt[i] < 0 || src_bt[i] >= src_n_blocks || dst_bt[i] < 0 || dst_bt[i] >= dst_n_blocks) fail(std::string(fn) + ": block id out of range");
  return static_cast<int64_t>(src_bt.size());
}

void launch_migration(int dev, const std::vector<int64_t>& src_bt, const std::vector<int64_t>& dst_bt, const void* src_ptr, void* dst_ptr, const Tensor& local_pool,
                      int64_t src_n_blocks, int64_t dst_n_blocks, int64_t layer_begin, int64_t layer_end) {
  const int64_t n = static_cast<int64_t>(src_bt.size());
  const int64_t run_bytes = local_pool.size(3) * local_pool.size(4) * local_pool.size(5) * local_pool.element_size();
  static const int64_t inline_blocks = hi_migrate_inline_table_blocks();
  if (n <= inline_blocks) {  // the tables ride in the kernel parameters: one launch, nothing in front of it
    int32_t tables[2][512];
    for (int64_t i = 0; i < n; ++i) {
      tables[0][i] = static_cast<int32_t>(src_bt[i]);
      tables[1][i] = static_cast<int32_t>(dst_bt[i]);
    }
    const HiPoolGeom src{local_pool.size(0), local_pool.size(1), src_n_blocks, run_bytes};
    const HiPoolGeom dst{local_pool.size(0), local_pool.size(1), dst_n_blocks, run_bytes};
    check(hi_migrate_blocks_host_tables(tables[0], tables[1], n, src_ptr, dst_ptr, src, dst, layer_begin, layer_end, dev, current_stream(dev)));
    return;
  }
  // both tables in one pinned staging tensor and one asynchronous H2D copy; torch's caching allocators keep the pinned block
  // and the device block alive until the copy / the kernel have run on this stream
  Tensor host = at::empty({2, n}, at::TensorOptions().dtype(at::kInt).pinned_memory(true));
  int32_t* h = host.data_ptr<int32_t>();
  for (int64_t i = 0; i < n; ++i) {
    h[i] = static_cast<int32_t>(src_bt[i]);
    h[n + i] = static_cast<int32_t>(dst_bt[i]);
  }
  Tensor tables = host.to(local_pool.options().dtype(at::kInt), /*non_blocking=*/true);
  const HiPoolGeom src{local_pool.size(0), local_pool.size(1), src_n_blocks, run_bytes};
  const HiPoolGeom dst{local_pool.size(0), local_pool.size(1), dst_n_blocks, run_bytes};
  check(hi_migrate_blocks_layers(tables.data_ptr<int32_t>(), tables.data_ptr<int32_t>() + n, n, src_ptr, dst_ptr, src, dst, layer_begin, layer_end, dev,
                                 current_stream(dev)));
}

// migrate_blocks restricted to layers [layer_begin, layer_end) (extension).
void migrate_blocks_layers(const std::vector<int64_t>& src_block_table, const std::vector<int64_t>& dst_block_table, const std::vector<int64_t>& src_cache,
                           Tensor dst_cache, int64_t src_cache_n_blocks, int64_t layer_begin, int64_t layer_end) {
  const int dev = require_cuda({&dst_cache});
  if (dst_cache.dim() != 6 || !dst_cache.is_contiguous()) fail("migrate_blocks: dst_cache must be a contiguous 6-D pool");  // block_migration.cpp:202
  if (check_tables("migrate_blocks", src_block_table, dst_block_table, src_cache_n_blocks, dst_cache.size(2)) == 0) return;
  void* src_ptr = nullptr;
  check(resolve(src_cache, dev, &src_ptr));
  launch_migration(dev, src_block_table, dst_block_table, src_ptr, dst_cache.data_ptr(), dst_cache, src_cache_n_blocks, dst_cache.size(2), layer_begin, layer_end);
}

// Copy blocks src_block_table[i] -> dst_block_table[i] for every (layer, K/V) plane of the pools (block_migration.cpp:194-245).
void migrate_blocks(const std::vector<int64_t>& src_block_table, const std::vector<int64_t>& dst_block_table, const std::vector<int64_t>& src_cache,
                    Tensor dst_cache, int64_t src_cache_n_blocks) {
  migrate_blocks_layers(src_block_table, dst_block_table, src_cache, dst_cache, src_cache_n_blocks, 0, dst_cache.dim() == 6 ? dst_cache.size(0) : 0);
}

// The same copy issued by the SENDER (extension): src_cache is the local pool, dst_cache the IPC handle of the receiver's pool.
void push_blocks(const std::vector<int64_t>& src_block_table, const std::vector<int64_t>& dst_block_table, Tensor src_cache, const std::vector<int64_t>& dst_cache,
                 int64_t dst_cache_n_blocks, int64_t layer_begin, int64_t layer_end) {
  const int dev = require_cuda({&src_cache});
  if (src_cache.dim() != 6 || !src_cache.is_contiguous()) fail("push_blocks: src_cache must be a contiguous 6-D pool");
  if (check_tables("push_blocks", src_block_table, dst_block_table, src_cache.size(2), dst_cache_n_blocks) == 0) return;
  if (layer_end < 0) layer_end = src_cache.size(0);
  void* dst_ptr = nullptr;
  check(resolve(dst_cache, dev, &dst_ptr));
  launch_migration(dev, src_block_table, dst_block_table, src_cache.data_ptr(), dst_ptr, src_cache, src_cache.size(2), dst_cache_n_blocks, layer_begin, layer_end);
}

// Both pools local tensors of one geometry except n_blocks (extension): the pack / unpack step of the NCCL backend
// (hydrainfer_b200/memory/communication.py), which ships a request as ONE contiguous staging pool instead of the reference's
// n_blocks * n_layers * 2 P2POps (communication.py:65-74).
void copy_blocks(const std::vector<int64_t>& src_block_table, const std::vector<int64_t>& dst_block_table, Tensor src_cache, Tensor dst_cache) {
  const int dev = require_cuda({&src_cache, &dst_cache});
  if (src_cache.dim() != 6 || dst_cache.dim() != 6 || !src_cache.is_contiguous() || !dst_cache.is_contiguous())
    fail("copy_blocks: pools must be contiguous 6-D tensors");
  if (src_cache.scalar_type() != dst_cache.scalar_type() || src_cache.size(0) != dst_cache.size(0) || src_cache.size(1) != dst_cache.size(1) ||
      src_cache.size(3) != dst_cache.size(3) || src_cache.size(4) != dst_cache.size(4) || src_cache.size(5) != dst_cache.size(5))
    fail("copy_blocks: pools may differ in n_blocks only");
  if (check_tables("copy_blocks", src_block_table, dst_block_table, src_cache.size(2), dst_cache.size(2)) == 0) return;
  launch_migration(dev, src_block_table, dst_block_table, src_cache.data_ptr(), dst_cache.data_ptr(), dst_cache, src_cache.size(2), dst_cache.size(2), 0,
                   dst_cache.size(0));
}

}  // namespace

PYBIND11_MODULE(kv_cache_kernels, m) {
  m.doc() = "set_kv_cache (hydrainfer_b200, sm_100a)";
  m.def("set_kv_cache", &set_kv_cache);
}

PYBIND11_MODULE(cache_kernels, m) {
  m.doc() = "cache kernels (hydrainfer_b200, sm_100a)";
  m.def("set_image_cache", &set_image_cache);
  m.def("get_image_cache", &get_image_cache);
}

PYBIND11_MODULE(position_embedding, m) {
  m.doc() = "position_embedding kernels (hydrainfer_b200, sm_100a)";
  m.def("apply_rotary_pos_emb", &apply_rotary_pos_emb);
  m.def("rope_set_kv_cache", &rope_set_kv_cache, py::arg("query"), py::arg("key"), py::arg("value"), py::arg("positions"), py::arg("cos_sin"),
        py::arg("rotary_dim"), py::arg("interleaved"), py::arg("slot_ids"), py::arg("key_cache"), py::arg("value_cache"), py::arg("write_back_k") = false,
        py::arg("force_scalar") = false);
}

PYBIND11_MODULE(flash_attn, m) {
  m.doc() = "paged / varlen attention (hydrainfer_b200, sm_100a: tcgen05 prefill, split-KV decode)";
  m.def("mha_varlen_fwd", &mha_varlen_fwd, py::arg("out"), py::arg("q"), py::arg("k"), py::arg("v"), py::arg("cu_seqlens_q"), py::arg("cu_seqlens_k"),
        py::arg("block_table_").none(true), py::arg("cu_block_lens").none(true), py::arg("alibi_slopes").none(true), py::arg("max_seqlen_q"),
        py::arg("max_seqlen_k"), py::arg("softmax_scale"), py::arg("softcap"), py::arg("window_size_left"), py::arg("window_size_right"),
        py::arg("num_splits"), py::arg("path") = 0, py::arg("work_items") = py::none(), py::arg("work_tile_tokens") = 0, py::arg("qk_work_hint") = 0);
  m.def("append_and_attend", &append_and_attend, py::arg("query"), py::arg("key"), py::arg("value"), py::arg("new_cache_slots"), py::arg("key_cache"),
        py::arg("value_cache"), py::arg("cu_seqlens_q"), py::arg("cu_seqlens_k"), py::arg("block_table"), py::arg("cu_block_lens"), py::arg("max_seqlen_q"),
        py::arg("max_seqlen_k"), py::arg("softmax_scale"), py::arg("path") = 0, py::arg("work_items") = py::none(), py::arg("work_tile_tokens") = 0,
        py::arg("qk_work_hint") = 0);
  m.def("last_launch_count", &last_launch_count);
  m.def("workspace", &workspace, py::arg("device"), py::arg("min_bytes") = 0);
}

PYBIND11_MODULE(block_migration, m) {
  m.doc() = "kv cache block migration (hydrainfer_b200, sm_100a)";
  m.def("get_ipc_mem_handle", &get_ipc_mem_handle);
  m.def("register_ipc_mem_handle", &register_ipc_mem_handle);
  m.def("migrate_blocks", &migrate_blocks);
  m.def("migrate_blocks_layers", &migrate_blocks_layers);
  m.def("push_blocks", &push_blocks, py::arg("src_block_table"), py::arg("dst_block_table"), py::arg("src_cache"), py::arg("dst_cache"),
        py::arg("dst_cache_n_blocks"), py::arg("layer_begin") = 0, py::arg("layer_end") = -1);
  m.def("copy_blocks", &copy_blocks);
  m.def("set_max_ctas", [](int64_t n) { check(hi_migrate_set_max_ctas(static_cast<int>(n))); }, py::arg("max_ctas"));
}
