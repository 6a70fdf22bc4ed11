// Microbenchmark (dev tool, GPU box): cycles per tcgen05.mma for the shapes the attention kernels issue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hydrainfer_b200/csrc -o /tmp/mma_probe tools/probes/mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace hi;

// mode 0: SS M128 N128 K16   1: SS M128 N64 K16   2: TS M128 N128 K16 (A from TMEM)   3: SS M128 N256   4: TS N64
template <int mode>
__global__ void __launch_bounds__(128, 1) probe(int reps, int same_desc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t barv;
  const uint32_t bar = ptx::smem_u32(&barv);
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&tmem_ptr), 512);
  ptx::tc_fence_before_sync(); __syncthreads(); ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x < 32) {
    const int n = mode == 1 || mode == 4 ? 64 : (mode == 3 ? 256 : 128);
    const uint32_t idesc = ptx::make_idesc_f16(true, false, false, 128, n);
    const uint64_t da = ptx::make_smem_desc_sw128(smem_base, 16, 1024);
    const uint64_t db = ptx::make_smem_desc_sw128(smem_base + 65536, 16, 1024);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (ptx::elect_one()) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t off = same_desc ? 0 : static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
          if (mode == 2 || mode == 4) ptx::mma_f16_ts(tmem + 256, tmem + kk * 8, db + off, idesc, true);
          else if (mode == 5) {  // two independent accumulators, alternating
            ptx::mma_f16_ss(tmem + (kk & 1) * 128, da + off + (kk & 1) * 2048, db + off, idesc, true);
          } else if (mode == 6) {  // SS (QK-like) and TS (PV-like) interleaved, independent accumulators
            if (kk & 1) ptx::mma_f16_ts(tmem + 256, tmem + 384 + (kk >> 1) * 8, db + off, idesc, true);
            else ptx::mma_f16_ss(tmem, da + off, db + off, idesc, true);
          } else if (mode == 7) {  // four independent SS accumulators round-robin
            ptx::mma_f16_ss(tmem + (kk & 3) * 128, da + off, db + off, idesc, true);
          }
          else ptx::mma_f16_ss(tmem, da + off, db + off, idesc, true);
        }
      }
      t1 = clock64();
      ptx::mma_commit(bar);
    }
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    t2 = clock64();
    if (ptx::elect_one() && blockIdx.x == 0) { out[0] = t1 - t0; }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = t2 - t0;
  }
  ptx::tc_fence_before_sync(); __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 512); }
}

template <int mode>
void run(long long* out, const char** names) {
  const int reps = 64;
  cudaFuncSetAttribute(probe<mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int it = 0; it < 2; ++it) { probe<mode><<<148, 128, 200 * 1024>>>(reps, 0, out); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return; } }
  printf("%-18s issue %.1f cyc/mma  complete %.1f cyc/mma\n", names[mode], (double)out[0] / (reps * 8), (double)out[1] / (reps * 8));
}

int main() {
  long long* out; cudaMallocManaged(&out, 64);
  const char* names[] = {"SS M128 N128 K16", "SS M128 N64 K16", "TS M128 N128 K16", "SS M128 N256 K16", "TS M128 N64 K16", "SS N128 2 accum", "SS/TS interleaved", "SS N128 4 accum"};
  run<0>(out, names); run<1>(out, names); run<2>(out, names); run<3>(out, names); run<4>(out, names); run<5>(out, names); run<6>(out, names); run<7>(out, names);
  return 0;
}
