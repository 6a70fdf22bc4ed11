// Microbenchmark (dev tool, GPU box): latency of the mbarrier hand-offs the attention kernels are built from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hydrainfer_b200/csrc -o tools/probes/mbar_probe.bin tools/probes/mbar_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace hi;

__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
template <int MODE>
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  if (MODE == 0) { while (!ptx::mbar_try_wait(bar, parity)) {} }
  else { while (!test_wait(bar, parity)) {} }
}

// mode 0/1: ping-pong between warp 0 (all 32 lanes arrive) and warp 1 (all 32 lanes arrive), try_wait / test_wait
// mode 2/3: warp 0..3 (128 threads) arrive on A; warp 4 waits A, one lane arrives on B (count 1); warps 0..3 wait B
// mode 4: one lane: tcgen05.mma (N128 SS) + commit, wait: latency of an MMA + commit round trip
// mode 5: one lane: commit only (no MMA pending), wait
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t barv[2];
  const uint32_t bar_a = ptx::smem_u32(&barv[0]), bar_b = ptx::smem_u32(&barv[1]);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_a, MODE >= 4 ? 1 : (MODE >= 2 ? 128 : 32));
    ptx::mbar_init(bar_b, MODE >= 2 ? 1 : 32);
    ptx::fence_mbar_init();
  }
  if (warp == 7) ptx::tmem_alloc(ptx::smem_u32(&tmem_ptr), 512);
  ptx::tc_fence_before_sync(); __syncthreads(); ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  const long long t0 = clock64();
  if (MODE <= 1) {
    if (warp == 0) { for (int r = 0; r < reps; ++r) { ptx::mbar_arrive(bar_a); wait<MODE & 1>(bar_b, r & 1); } }
    if (warp == 1) { for (int r = 0; r < reps; ++r) { wait<MODE & 1>(bar_a, r & 1); ptx::mbar_arrive(bar_b); } }
  } else if (MODE <= 3) {
    if (warp < 4) { for (int r = 0; r < reps; ++r) { ptx::mbar_arrive(bar_a); wait<MODE & 1>(bar_b, r & 1); } }
    if (warp == 4) { for (int r = 0; r < reps; ++r) { wait<MODE & 1>(bar_a, r & 1); if (ptx::elect_one()) ptx::mbar_arrive(bar_b); __syncwarp(); } }
  } else {
    if (warp == 0) {
      const uint32_t idesc = ptx::make_idesc_f16(true, false, false, 128, 128);
      const uint64_t da = ptx::make_smem_desc_sw128(smem_base, 16, 1024), db = ptx::make_smem_desc_sw128(smem_base + 65536, 16, 1024);
      for (int r = 0; r < reps; ++r) {
        if (ptx::elect_one()) { if (MODE == 4) ptx::mma_f16_ss(tmem, da, db, idesc, false); ptx::mma_commit(bar_a); }
        __syncwarp();
        wait<0>(bar_a, r & 1);
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  ptx::tc_fence_before_sync(); __syncthreads();
  if (warp == 7) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 512); }
}

template <int MODE>
void run(long long* out, const char* name) {
  const int reps = 1000;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int it = 0; it < 2; ++it) { probe<MODE><<<148, 256, 200 * 1024>>>(reps, out); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return; } }
  printf("%-60s %.1f cycles per round trip\n", name, (double)out[0] / reps);
}

int main() {
  long long* out; cudaMallocManaged(&out, 64);
  run<0>(out, "warp<->warp ping-pong, try_wait");
  run<1>(out, "warp<->warp ping-pong, test_wait spin");
  run<2>(out, "128 threads -> warp -> 128 threads, try_wait");
  run<3>(out, "128 threads -> warp -> 128 threads, test_wait spin");
  run<4>(out, "one MMA (M128 N128 K16) + commit + wait");
  run<5>(out, "commit (nothing pending) + wait");
  return 0;
}
