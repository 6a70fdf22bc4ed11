// KV-page migration: one gather kernel over (possibly peer-mapped) pools + cached CUDA-IPC mappings.
//
// Replaces csrc/data_transfer/block_migration.cpp of the reference: migrate_blocks (:194-245) issues
// n_layers * n_tokens * n_blocks cudaMemcpyAsync calls (262 144 for a 4096-block LLaVA-7B request) and
// re-opens the IPC handle on every call (:213-215).  Here the peer pool is mapped once per process and all
// runs of a request move in ONE launch: the destination GPU pulls through NVLink with 128-bit loads and
// writes its own HBM, the same direction the reference's receiver-side memcpy moves data.
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace hi {

struct MigrateArgs {
  const int32_t* src_blocks;
  const int32_t* dst_blocks;
  const char* src_pool;
  char* dst_pool;
  int64_t n;               // blocks to move
  int64_t planes;          // n_layers * n_tokens
  int64_t run_bytes;       // contiguous bytes of one (plane, block)
  int64_t src_plane_bytes; // src n_blocks * run_bytes
  int64_t dst_plane_bytes;
  int64_t plane_begin;     // first (layer, kv) plane moved by this launch
  int64_t pieces_per_run;  // run_bytes / kPieceBytes (rounded up)
  int64_t total_pieces;
  int bulk_stages;         // bulk-copy variant: 16-KiB shared-memory stages per CTA
};

constexpr int kMigrateThreads = 256;
constexpr int kVecPerThread = 4;                                       // independent 16-B loads in flight per thread
constexpr int64_t kPieceBytes = kMigrateThreads * kVecPerThread * 16;  // 16 KiB handled by a CTA per step

// Persistent CTAs walk 16-KiB pieces of the (plane, block) runs.  A run is contiguous in both pools
// (INDEX_6D of block_migration.cpp:26-27 with the three innermost indices 0), so every access is a
// fully coalesced 128-bit vector; each thread keeps kVecPerThread loads in flight to cover NVLink latency.
template <typename Tables>
__device__ __forceinline__ void migrate_gather_body(const MigrateArgs& a, const Tables& tables) {
  for (int64_t piece = blockIdx.x; piece < a.total_pieces; piece += gridDim.x) {
    const int64_t run = piece / a.pieces_per_run;
    const int64_t off = (piece - run * a.pieces_per_run) * kPieceBytes;
    const int64_t plane_rel = run / a.n;
    const int64_t i = run - plane_rel * a.n;
    const int64_t plane = a.plane_begin + plane_rel;
    const int64_t sb = tables.src(i);
    const int64_t db = tables.dst(i);
    const char* __restrict__ src = a.src_pool + plane * a.src_plane_bytes + sb * a.run_bytes + off;
    char* __restrict__ dst = a.dst_pool + plane * a.dst_plane_bytes + db * a.run_bytes + off;
    const int64_t left = a.run_bytes - off;
    const int n_vec = static_cast<int>((left < kPieceBytes ? left : kPieceBytes) >> 4);
    uint4 v[kVecPerThread];
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      const int idx = threadIdx.x + k * kMigrateThreads;
      if (idx < n_vec) v[k] = ldg_stream_16(src + static_cast<int64_t>(idx) * 16);
    }
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      const int idx = threadIdx.x + k * kMigrateThreads;
      if (idx < n_vec) *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(idx) * 16) = v[k];
    }
  }
}

struct DeviceTables {
  const int32_t* s;
  const int32_t* d;
  __device__ __forceinline__ int64_t src(int64_t i) const { return __ldg(s + i); }
  __device__ __forceinline__ int64_t dst(int64_t i) const { return __ldg(d + i); }
};
__global__ void __launch_bounds__(kMigrateThreads) migrate_gather_kernel(const MigrateArgs a) {
  migrate_gather_body(a, DeviceTables{a.src_blocks, a.dst_blocks});
}

// Small requests carry their block tables IN the kernel parameters: no staging buffer, no host-to-device copy in front of the
// launch (a 16-block request is a few tens of microseconds of transfer, comparable with an extra stream operation).
constexpr int kInlineBlocks = 384;  // 2 x 384 x 4 B = 3 KiB of the 4 KiB parameter space
struct MigrateInlineArgs {
  MigrateArgs base;
  int32_t src[kInlineBlocks];
  int32_t dst[kInlineBlocks];
};
struct InlineTables {
  const MigrateInlineArgs* p;
  __device__ __forceinline__ int64_t src(int64_t i) const { return p->src[i]; }
  __device__ __forceinline__ int64_t dst(int64_t i) const { return p->dst[i]; }
};
__global__ void __launch_bounds__(kMigrateThreads) migrate_gather_inline_kernel(const __grid_constant__ MigrateInlineArgs a) {
  migrate_gather_body(a.base, InlineTables{&a});
}

// ---- bulk-copy variant: the TMA engine moves the bytes ---------------------------------------------------------------------
// For LARGE transfers from / to peer memory.  One warp per CTA, one CTA per SM: an elected lane issues a 1-D bulk copy
// (cp.async.bulk) of each 16-KiB piece from the source pool into a shared-memory stage and, when its bytes have landed, a bulk
// copy from the stage to the destination pool; two stages (32 KiB) keep two pieces in flight per SM = 4.7 MB over the GPU, twice
// what 800 GB/s x 3 us of NVLink latency needs.  No thread touches the data, the SM's issue slots stay with whatever else runs
// on the GPU (the decode step of the receiving instance, epdnode.py:362-447), and 32 KiB of shared memory fit beside three
// resident decode CTAs of 64 KiB.
constexpr int kBulkStages = 2;      // default; HI_MIGRATE_BULK_STAGES (2..8) for experiments
constexpr int kBulkMaxStages = 8;
constexpr int kBulkThreads = 32;
constexpr uint32_t kBulkPiece = 16384;
struct BulkPiece {
  const char* src;
  char* dst;
  uint32_t bytes;
};
template <typename Tables>
__device__ __forceinline__ BulkPiece bulk_piece(const MigrateArgs& a, const Tables& tables, int64_t piece) {
  const int64_t run = piece / a.pieces_per_run;
  const int64_t off = (piece - run * a.pieces_per_run) * kBulkPiece;
  const int64_t plane_rel = run / a.n;
  const int64_t i = run - plane_rel * a.n;
  const int64_t plane = a.plane_begin + plane_rel;
  const int64_t left = a.run_bytes - off;
  BulkPiece p;
  p.src = a.src_pool + plane * a.src_plane_bytes + tables.src(i) * a.run_bytes + off;
  p.dst = a.dst_pool + plane * a.dst_plane_bytes + tables.dst(i) * a.run_bytes + off;
  p.bytes = static_cast<uint32_t>(left < kBulkPiece ? left : kBulkPiece);
  return p;
}
template <typename Tables>
__device__ __forceinline__ void migrate_bulk_body(const MigrateArgs& a, const Tables& tables) {
  extern __shared__ __align__(128) uint8_t bulk_smem[];
  __shared__ __align__(8) uint64_t full[kBulkMaxStages];
  const uint32_t stage0 = ptx::smem_u32(bulk_smem);
  const int n_st = a.bulk_stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_st; ++s) ptx::mbar_init(ptx::smem_u32(&full[s]), 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  // pieces blockIdx.x, blockIdx.x + gridDim.x, ...: local index k.  Destination and size of the pieces in flight live in shared memory.
  __shared__ BulkPiece ring[kBulkMaxStages];
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t n_local = first < a.total_pieces ? (a.total_pieces - first + stride - 1) / stride : 0;
  for (int s = 0; s < n_st && s < n_local; ++s) {
    const BulkPiece p = bulk_piece(a, tables, first + s * stride);
    ring[s] = p;
    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full[s]), p.bytes);
    ptx::bulk_load_1d(stage0 + s * kBulkPiece, p.src, p.bytes, ptx::smem_u32(&full[s]));
  }
  int s = 0;
  uint32_t phase = 0;
  for (int64_t k = 0; k < n_local; ++k) {
    const bool more = k + n_st < n_local;
    BulkPiece nxt{};
    if (more) nxt = bulk_piece(a, tables, first + (k + n_st) * stride);  // table loads in flight behind the wait below
    ptx::mbar_wait(ptx::smem_u32(&full[s]), phase);
    const BulkPiece p = ring[s];
    ptx::bulk_store_1d(p.dst, stage0 + s * kBulkPiece, p.bytes);
    ptx::bulk_commit_group();
    if (more) {
      ptx::bulk_wait_group_read<0>();  // the store has read the stage: it may be refilled
      ring[s] = nxt;
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full[s]), nxt.bytes);
      ptx::bulk_load_1d(stage0 + s * kBulkPiece, nxt.src, nxt.bytes, ptx::smem_u32(&full[s]));
    }
    if (++s == n_st) {
      s = 0;
      phase ^= 1u;
    }
  }
  ptx::bulk_wait_group<0>();  // the writes are done before the kernel ends
}
__global__ void __launch_bounds__(kBulkThreads) migrate_bulk_kernel(const MigrateArgs a) {
  migrate_bulk_body(a, DeviceTables{a.src_blocks, a.dst_blocks});
}
__global__ void __launch_bounds__(kBulkThreads) migrate_bulk_inline_kernel(const __grid_constant__ MigrateInlineArgs a) {
  migrate_bulk_body(a.base, InlineTables{&a});
}

// ---- IPC mapping cache -------------------------------------------------------------------------------------------
struct IpcEntry {
  uint8_t handle[64];
  int device;
  void* base;
  size_t size;  // bytes of the mapped allocation (0 if the driver would not tell)
};
static std::mutex g_ipc_mu;
static std::vector<IpcEntry> g_ipc_entries;

// A pool another process exported (cudaPointerGetAttributes reports the MAPPING device for such memory, not the exporting GPU).
static bool pointer_is_ipc_mapped(const void* ptr) {
  std::lock_guard<std::mutex> lock(g_ipc_mu);
  for (const IpcEntry& e : g_ipc_entries) {
    const char* b = static_cast<const char*>(e.base);
    if (ptr >= b && (e.size == 0 ? ptr == b : ptr < b + e.size)) return true;
  }
  return false;
}

}  // namespace hi

extern "C" int hi_migrate_blocks(const int32_t* src_blocks, const int32_t* dst_blocks, int64_t n, const void* src_pool,
                                 void* dst_pool, HiPoolGeom src, HiPoolGeom dst, int device, void* stream) {
  return hi_migrate_blocks_layers(src_blocks, dst_blocks, n, src_pool, dst_pool, src, dst, 0, src.n_layers, device, stream);
}

namespace hi {
static std::atomic<int> g_migrate_max_ctas{0};  // 0 = no cap (process-wide tuning knob)

struct IpcEntry;
static bool pointer_is_ipc_mapped(const void* ptr);  // defined behind the IPC cache below

// Device that owns the allocation behind `ptr` (peer-mapped and IPC-mapped pools report the exporting GPU); `fallback` when the
// runtime cannot tell.  The last few answers are cached: pools live for the life of the process.
static int pointer_device(const void* ptr, int fallback) {
  struct Entry {
    const void* ptr;
    int device;
  };
  static std::mutex mu;
  static Entry cache[8] = {};
  static int next = 0;
  std::lock_guard<std::mutex> lock(mu);
  for (const Entry& e : cache)
    if (e.ptr == ptr && ptr != nullptr) return e.device;
  cudaPointerAttributes attr;
  int dev = fallback;
  if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice) dev = attr.device;
  else (void)cudaGetLastError();
  cache[next] = Entry{ptr, dev};
  next = (next + 1) % 8;
  return dev;
}
// Shared by the device-table and the inline-table entry points: checks + geometry -> MigrateArgs and the grid size.
static int prepare_migration(MigrateArgs& a, int64_t& grid, bool& bulk, int64_t n, const void* src_pool, void* dst_pool, const HiPoolGeom& src,
                             const HiPoolGeom& dst, int64_t layer_begin, int64_t layer_end, int device) {
  bulk = false;
  HI_CHECK_ARG(src_pool && dst_pool, "migrate_blocks: null pointer");
  HI_CHECK_ARG(src.n_layers == dst.n_layers && src.n_tokens == dst.n_tokens && src.run_bytes == dst.run_bytes,
               "migrate_blocks: pools differ in more than n_blocks (layers %lld/%lld, tokens %lld/%lld, run bytes %lld/%lld)",
               (long long)src.n_layers, (long long)dst.n_layers, (long long)src.n_tokens, (long long)dst.n_tokens,
               (long long)src.run_bytes, (long long)dst.run_bytes);
  HI_CHECK_ARG(src.run_bytes > 0 && src.run_bytes % 16 == 0, "migrate_blocks: run of %lld bytes is not a multiple of 16",
               (long long)src.run_bytes);
  HI_CHECK_ARG(aligned_to(src_pool, 16) && aligned_to(dst_pool, 16), "migrate_blocks: pools must be 16-byte aligned");
  a.src_pool = static_cast<const char*>(src_pool);
  a.dst_pool = static_cast<char*>(dst_pool);
  a.n = n;
  a.planes = (layer_end - layer_begin) * src.n_tokens;
  a.plane_begin = layer_begin * src.n_tokens;
  a.run_bytes = src.run_bytes;
  a.src_plane_bytes = src.n_blocks * src.run_bytes;
  a.dst_plane_bytes = dst.n_blocks * dst.run_bytes;
  a.pieces_per_run = (a.run_bytes + kPieceBytes - 1) / kPieceBytes;
  a.total_pieces = a.planes * n * a.pieces_per_run;
  const int sm_count = sm_count_of(device);
  HI_CHECK_ARG(sm_count > 0, "migrate_blocks: cannot read the SM count of device %d", device);
  // 8 resident CTAs of 256 threads per SM = 128 KiB of loads in flight per SM: what a pool -> pool copy inside one GPU's HBM wants
  // (2.9 TB/s), and what a small request wants (all its pieces in flight in one round trip).  A LARGE transfer over NVLink needs
  // far less - 800 GB/s x 3 us = 2.4 MB in flight, 16 KiB per SM - and 8 CTAs per SM take every thread slot of the receiving
  // GPU away from the decode step running beside the pull.  bench.py migrate_under_decode on 2 B200 (pull alone / pull under
  // decode / slowdown of the decode step): 8 CTAs per SM 778 / 769 GB/s / 1.61x, 2 per SM 773 / 463 / 1.52x, 1 per SM
  // 761 / 555 / 1.33x, 64 CTAs 410 / 255 / 1.19x.  One CTA per SM keeps 98 % of the rate alone and is the best trade under load.
  // hi_migrate_set_max_ctas() overrides the choice.
  grid = static_cast<int64_t>(sm_count) * 8;
  const int cap = g_migrate_max_ctas.load(std::memory_order_relaxed);
  if (cap > 0) {
    if (grid > cap) grid = cap;
  } else if (pointer_is_ipc_mapped(src_pool) || pointer_is_ipc_mapped(dst_pool) || pointer_device(src_pool, device) != device ||
             pointer_device(dst_pool, device) != device) {
    // ... and the TMA engine can move them (migrate_bulk_kernel, one warp per SM); small requests included - 16-block requests run
    // 712 / 443 GB/s (LLaVA-7B / Qwen2-VL-7B pools) against 694-707 / 408-441 with the load / store kernel.  HI_MIGRATE_BULK=0 keeps
    // the load / store kernel (one CTA per SM for large transfers, as measured above).
    const char* env = tuning_env("HI_MIGRATE_BULK");
    bulk = !(env != nullptr && env[0] == '0');
    if (bulk || a.total_pieces > grid * 4) grid = static_cast<int64_t>(sm_count);
    a.bulk_stages = kBulkStages;
    if (const char* st = tuning_env("HI_MIGRATE_BULK_STAGES")) a.bulk_stages = atoi(st) < 2 ? 2 : atoi(st) > kBulkMaxStages ? kBulkMaxStages : atoi(st);
  }
  if (const char* env = tuning_env("HI_MIGRATE_BULK")) {
    if (env[0] == '2') {  // test override: the bulk-copy kernel for EVERY launch (a one-GPU box covers it that way)
      bulk = true;
      a.bulk_stages = kBulkStages;
      if (cap <= 0) grid = static_cast<int64_t>(sm_count);
    }
  }
  if (grid > a.total_pieces) grid = a.total_pieces;
  return HI_OK;
}
}  // namespace hi

extern "C" int hi_migrate_set_max_ctas(int max_ctas) {
  hi::g_migrate_max_ctas.store(max_ctas > 0 ? max_ctas : 0, std::memory_order_relaxed);
  return HI_OK;
}

extern "C" int hi_migrate_blocks_layers(const int32_t* src_blocks, const int32_t* dst_blocks, int64_t n, const void* src_pool,
                                        void* dst_pool, HiPoolGeom src, HiPoolGeom dst, int64_t layer_begin, int64_t layer_end,
                                        int device, void* stream) {
  using namespace hi;
  reset_launch_count();
  HI_CHECK_ARG(n >= 0, "migrate_blocks: negative block count");
  HI_CHECK_ARG(layer_begin >= 0 && layer_begin <= layer_end && layer_end <= src.n_layers,
               "migrate_blocks: layer range [%lld, %lld) outside the pool's %lld layers", (long long)layer_begin,
               (long long)layer_end, (long long)src.n_layers);
  if (n == 0 || layer_begin == layer_end) return HI_OK;
  HI_CHECK_ARG(src_blocks && dst_blocks, "migrate_blocks: null pointer");
  MigrateArgs a{};
  int64_t grid = 0;
  bool bulk = false;
  const int rc = prepare_migration(a, grid, bulk, n, src_pool, dst_pool, src, dst, layer_begin, layer_end, device);
  if (rc != HI_OK) return rc;
  a.src_blocks = src_blocks;
  a.dst_blocks = dst_blocks;
  HI_DEVICE_GUARD(device);
  if (bulk) {
    static PerDeviceFlags configured;
    HI_CUDA(configure_dynamic_smem(configured, migrate_bulk_kernel, kBulkMaxStages * kBulkPiece));
    migrate_bulk_kernel<<<static_cast<unsigned>(grid), kBulkThreads, a.bulk_stages * kBulkPiece, static_cast<cudaStream_t>(stream)>>>(a);
  }
  else migrate_gather_kernel<<<static_cast<unsigned>(grid), kMigrateThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

extern "C" int hi_migrate_blocks_host_tables(const int32_t* src_blocks_host, const int32_t* dst_blocks_host, int64_t n, const void* src_pool,
                                             void* dst_pool, HiPoolGeom src, HiPoolGeom dst, int64_t layer_begin, int64_t layer_end,
                                             int device, void* stream) {
  using namespace hi;
  reset_launch_count();
  HI_CHECK_ARG(n >= 0 && n <= kInlineBlocks, "migrate_blocks: %lld blocks exceed the %d an inline table holds (use hi_migrate_blocks_layers)",
               (long long)n, kInlineBlocks);
  HI_CHECK_ARG(layer_begin >= 0 && layer_begin <= layer_end && layer_end <= src.n_layers,
               "migrate_blocks: layer range [%lld, %lld) outside the pool's %lld layers", (long long)layer_begin,
               (long long)layer_end, (long long)src.n_layers);
  if (n == 0 || layer_begin == layer_end) return HI_OK;
  HI_CHECK_ARG(src_blocks_host && dst_blocks_host, "migrate_blocks: null pointer");
  MigrateInlineArgs a{};
  int64_t grid = 0;
  bool bulk = false;
  const int rc = prepare_migration(a.base, grid, bulk, n, src_pool, dst_pool, src, dst, layer_begin, layer_end, device);
  if (rc != HI_OK) return rc;
  std::memcpy(a.src, src_blocks_host, static_cast<size_t>(n) * sizeof(int32_t));
  std::memcpy(a.dst, dst_blocks_host, static_cast<size_t>(n) * sizeof(int32_t));
  HI_DEVICE_GUARD(device);
  if (bulk) {
    static PerDeviceFlags configured;
    HI_CUDA(configure_dynamic_smem(configured, migrate_bulk_inline_kernel, kBulkMaxStages * kBulkPiece));
    migrate_bulk_inline_kernel<<<static_cast<unsigned>(grid), kBulkThreads, a.base.bulk_stages * kBulkPiece, static_cast<cudaStream_t>(stream)>>>(a);
  }
  else migrate_gather_inline_kernel<<<static_cast<unsigned>(grid), kMigrateThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  note_launch();
  HI_CUDA(cudaGetLastError());
  return HI_OK;
}

extern "C" int hi_migrate_inline_table_blocks(void) { return hi::kInlineBlocks; }

extern "C" int hi_ipc_get_handle(const void* ptr, uint8_t handle_out[64], int64_t* offset_out, int device) {
  using namespace hi;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  HI_CHECK_ARG(ptr && handle_out && offset_out, "ipc_get_handle: null pointer");
  HI_DEVICE_GUARD(device);
  cudaIpcMemHandle_t h;
  HI_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  std::memcpy(handle_out, &h, 64);
  // The handle names the whole allocation; recover where ptr sits inside it.
  void* base = nullptr;
  size_t size = 0;
  typedef int (*GetRangeFn)(void**, size_t*, void*);  // cuMemGetAddressRange(CUdeviceptr*, size_t*, CUdeviceptr)
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  HI_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("ipc_get_handle: cuMemGetAddressRange not available");
    return HI_ERR_CUDA;
  }
  const int rc = reinterpret_cast<GetRangeFn>(fn)(&base, &size, const_cast<void*>(ptr));
  if (rc != 0) {
    set_error("ipc_get_handle: cuMemGetAddressRange failed with CUresult %d", rc);
    return HI_ERR_CUDA;
  }
  *offset_out = static_cast<const char*>(ptr) - static_cast<const char*>(base);
  return HI_OK;
}

extern "C" int hi_ipc_open_handle(const uint8_t handle[64], int64_t offset, int device, void** ptr_out) {
  using namespace hi;
  HI_CHECK_ARG(handle && ptr_out && offset >= 0, "ipc_open_handle: bad argument");
  std::lock_guard<std::mutex> lock(g_ipc_mu);
  for (const IpcEntry& e : g_ipc_entries) {
    if (e.device == device && std::memcmp(e.handle, handle, 64) == 0) {
      *ptr_out = static_cast<char*>(e.base) + offset;
      return HI_OK;
    }
  }
  HI_DEVICE_GUARD(device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  void* base = nullptr;
  const cudaError_t err = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (err == cudaErrorPeerAccessUnsupported) {
    (void)cudaGetLastError();
    set_error("ipc_open_handle: peer access to the exporting GPU is unsupported from device %d", device);
    return HI_ERR_PEER_UNSUPPORTED;
  }
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(err));
    return HI_ERR_CUDA;
  }
  IpcEntry e;
  std::memcpy(e.handle, handle, 64);
  e.device = device;
  e.base = base;
  e.size = 0;
  {
    typedef int (*GetRangeFn)(void**, size_t*, void*);  // cuMemGetAddressRange
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    void* rbase = nullptr;
    size_t rsize = 0;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn != nullptr &&
        qres == cudaDriverEntryPointSuccess && reinterpret_cast<GetRangeFn>(fn)(&rbase, &rsize, base) == 0)
      e.size = rsize;
    else
      (void)cudaGetLastError();
  }
  g_ipc_entries.push_back(e);
  *ptr_out = static_cast<char*>(base) + offset;
  return HI_OK;
}

extern "C" int hi_ipc_close_all(void) {
  using namespace hi;
  std::lock_guard<std::mutex> lock(g_ipc_mu);
  int rc = HI_OK;
  for (const IpcEntry& e : g_ipc_entries) {
    DeviceGuard guard(e.device);
    if (!guard.ok || cudaIpcCloseMemHandle(e.base) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("ipc_close_all: cudaIpcCloseMemHandle failed");
      rc = HI_ERR_CUDA;
    }
  }
  g_ipc_entries.clear();
  return rc;
}

extern "C" int hi_peer_copy(void* dst, int dst_device, const void* src, int src_device, int64_t bytes, void* stream) {
  using namespace hi;
  HI_CHECK_ARG(dst && src && bytes >= 0, "peer_copy: bad argument");
  HI_CUDA(cudaMemcpyPeerAsync(dst, dst_device, src, src_device, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream)));
  return HI_OK;
}

extern "C" int hi_enable_peer_access(int device, int peer_device) {
  using namespace hi;
  if (device == peer_device) return HI_OK;
  int can = 0;
  HI_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
  if (!can) {
    set_error("enable_peer_access: device %d cannot access device %d", device, peer_device);
    return HI_ERR_PEER_UNSUPPORTED;
  }
  HI_DEVICE_GUARD(device);
  const cudaError_t err = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (err == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return HI_OK;
  }
  HI_CUDA(err);
  return HI_OK;
}
