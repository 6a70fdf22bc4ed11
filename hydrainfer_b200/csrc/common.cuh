// Shared host/device helpers for the hi_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../include/hi_b200.h"

namespace hi {

// Tuning / test overrides come from the environment; an empty value means "not set" (HI_X= in a shell loop must not turn
// into atoi("") == 0).
inline const char* tuning_env(const char* name) {
  const char* e = getenv(name);
  return (e != nullptr && e[0] != '\0') ? e : nullptr;
}

// ---- error plumbing -------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void note_launch();          // counts kernels launched by the current API call
void reset_launch_count();
// bench hooks: record the armed event pair around the dominant kernel of the current call (no-ops when disarmed)
void timing_mark_start(cudaStream_t stream);
void timing_mark_stop(cudaStream_t stream);

#define HI_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::hi::set_error(__VA_ARGS__);        \
      return HI_ERR_INVALID_ARGUMENT;      \
    }                                      \
  } while (0)

#define HI_CHECK_SUPPORTED(cond, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::hi::set_error(__VA_ARGS__);        \
      return HI_ERR_UNSUPPORTED;           \
    }                                      \
  } while (0)

#define HI_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if (err__ != cudaSuccess) {                                                                    \
      ::hi::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(err__)); \
      return HI_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

// ---- per-device state ------------------------------------------------------------------------------------------------------
// One process may drive several GPUs (same-process peer migration, tests, a multi-GPU engine) and several threads (the
// engine thread and the image-embed thread): everything that is per device / per context lives in arrays indexed
// by the device ordinal, written with atomics (racing writers store the same value).
constexpr int kMaxDevices = 64;

// Makes `device` current for the duration of a C-ABI call and restores the caller's current device on the way out, so
// a launch on cuda:1 never changes what torch (or any other runtime user on this thread) sees as the current device.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
    if (cur != device) {
      ok = device >= 0 && device < kMaxDevices && cudaSetDevice(device) == cudaSuccess;
      prev = cur;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define HI_DEVICE_GUARD(device)                                              \
  ::hi::DeviceGuard device_guard__(device);                                  \
  do {                                                                       \
    if (!device_guard__.ok) {                                                \
      ::hi::set_error("cudaSetDevice(%d) failed: %s", static_cast<int>(device), cudaGetErrorString(cudaGetLastError())); \
      return HI_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: call it once per (kernel instance, device).
// `flags` is a function-local static of the launching template instance.
struct PerDeviceFlags {
  std::atomic<unsigned char> done[kMaxDevices];
};
template <typename Kernel>
inline cudaError_t configure_dynamic_smem(PerDeviceFlags& flags, Kernel kernel, int bytes) {
  int device = 0;
  cudaError_t err = cudaGetDevice(&device);
  if (err != cudaSuccess) return err;
  if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
  if (flags.done[device].load(std::memory_order_acquire)) return cudaSuccess;
  err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err == cudaSuccess) flags.done[device].store(1, std::memory_order_release);
  return err;
}

// ---- split-KV scratch ---------------------------------------------------------------------------------------------------------
// Every kernel's split rule is a function of the launch extents only, never of the size of the caller's workspace: a launch
// whose partials (n_tokens * n_qo_heads * n_splits entries of head_dim + 2 floats) would exceed kMaxPartialBytes takes fewer
// splits (such launches have thousands of tiles and do not need them), and a workspace smaller than what the rule asks for
// is an error (HI_ERR_WORKSPACE), not a quiet change of the split count.  hi_attention_workspace_bytes() is the matching bound.
constexpr int64_t kMaxPartialBytes = int64_t(256) << 20;
constexpr int64_t kWorkspaceTailBytes = 512;  // work counter (last 256 bytes, 256-byte aligned) + slack
inline int64_t partial_bytes_per_split(int64_t n_tokens, int n_qo_heads, int head_dim) {
  return n_tokens * n_qo_heads * static_cast<int64_t>(head_dim + 2) * 4;
}
inline int cap_splits(int n_splits, int64_t n_tokens, int n_qo_heads, int head_dim) {
  const int64_t per_split = partial_bytes_per_split(n_tokens, n_qo_heads, head_dim);
  int64_t most = per_split > 0 ? kMaxPartialBytes / per_split : 1;
  if (most < 1) most = 1;
  return n_splits > most ? static_cast<int>(most) : n_splits;
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------------------
// The kernels of one layer call (append -> attention -> split merge) are short, and for small decode batches the gap between two
// dependent launches (grid launch + CTA scheduling + the prologue: barrier init, TMEM allocation, tensor-map prefetch) is a
// visible share of the call.  Every kernel of the path is therefore launched with programmatic stream serialization: it may
// start while its predecessor is still running, executes pdl_wait() before it touches anything a predecessor kernel produced
// (griddepcontrol.wait returns once the prerequisite grids have completed and their writes are visible), and calls
// pdl_launch_dependents() early so that ITS successor can do the same.  A predecessor that is not one of ours (it never triggers)
// simply completes first: the ordinary stream order.  HI_PDL=0 switches the launch attribute off (A/B, debugging).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* env = tuning_env("HI_PDL");
    return !(env != nullptr && env[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// SM count of `device`, cached per device (0 on error).
inline int sm_count_of(int device) {
  static std::atomic<int> cache[kMaxDevices];
  if (device < 0 || device >= kMaxDevices) return 0;
  int n = cache[device].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    cache[device].store(n, std::memory_order_relaxed);
  }
  return n;
}

inline int dtype_size(int dtype) {
  switch (dtype) {
    case HI_F32: return 4;
    case HI_F16: return 2;
    case HI_BF16: return 2;
    default: return 0;
  }
}

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- device: element conversion ---------------------------------------------------------------------------
template <typename T>
struct Elem;

template <>
struct Elem<float> {
  static constexpr int kBytes = 4;
  __device__ static __forceinline__ float to_f32(float v) { return v; }
  __device__ static __forceinline__ float from_f32(float v) { return v; }
};
template <>
struct Elem<__half> {
  static constexpr int kBytes = 2;
  __device__ static __forceinline__ float to_f32(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kBytes = 2;
  __device__ static __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
};

// Unpack one 32-bit word holding two 16-bit elements (low half = lower address).
template <typename T>
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi);
template <>
__device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t w, float& lo, float& hi) {
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xffff0000u);
}
template <>
__device__ __forceinline__ void unpack2<__half>(uint32_t w, float& lo, float& hi) {
  const __half2 h = *reinterpret_cast<const __half2*>(&w);
  const float2 f = __half22float2(h);
  lo = f.x;
  hi = f.y;
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Streaming 16-byte global load that does not pollute L1 (KV pages are read once per launch).
__device__ __forceinline__ uint4 ldg_stream_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream_8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sm_100 arithmetic helpers: 3-input max (FMNMX3) and packed fp32x2 FMA / ADD (FFMA2 / FADD2).
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t"
      ".reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// exp2 on the FMA / ALU pipes for two values at once (Cody-Waite range reduction + degree-3 minimax polynomial, max
// relative error 7.5e-5 - far below the 16-bit rounding P gets anyway).  Used for a fraction of the softmax exponentials
// so that MUFU.EX2 (16 lanes/clk/SM) stops being the only pipe that limits the softmax warps.  x <= ~100; x -> -inf gives
// 2^-126 (~1e-38) instead of 0.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 r = fadd2(x, make_float2(12582912.f, 12582912.f));     // 1.5 * 2^23: low mantissa bits of r = round(x)
  const float2 n = fadd2(r, make_float2(-12582912.f, -12582912.f));
  const float2 f = ffma2(n, make_float2(-1.f, -1.f), x);               // f in [-0.5, 0.5]
  float2 p = ffma2(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.2426111400f, 0.2426111400f));
  p = ffma2(p, f, make_float2(0.6932609677f, 0.6932609677f));
  p = ffma2(p, f, make_float2(0.9999280572f, 0.9999280572f));
  float2 y;
  y.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23));  // add round(x) to the exponent field
  y.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23));
  return y;
}

}  // namespace hi
