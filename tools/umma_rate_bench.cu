// Microbenchmark: issue rate of tcgen05.mma.kind::tf32 (M = 128, K = 8) from one thread, as a function of N, of the
// number of independent TMEM accumulators the MMAs rotate over, and of the shared-memory layout.  One CTA per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I kaldi-lstm_b200/csrc -o tools/_build/umma_rate_bench tools/umma_rate_bench.cu
#include <cstdio>
#include <vector>
#include "lstmp_tc.cuh"
using namespace lstmp;
using namespace lstmp::tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128, 1) rate(int N, int nacc, int swz, int nmma, int kind_f16, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 * 1024) / 4; i += 128) reinterpret_cast<float*>(tiles)[i] = 1.0f;
  if (tid == 0) {
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // mode 2: warp-uniform, the burst of 8 MMAs fully unrolled with static operand offsets (the CUTLASS pattern: every
  // MMA of a burst gets its own uniform registers)
  if (mode == 2 && warp == 0) {
    const uint32_t a = smem_u32(tiles), b = a + 32 * 1024;
    const uint32_t idesc = idesc_tf32(128, N);
    const int spacing = 512 / nacc, amask = nacc - 1;
    long long t0 = clock64();
    for (int i = 0; i < nmma; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint64_t da = make_desc_sw128(a + 32 * (u & 3) + 4096 * (u >> 2));
        const uint64_t db = make_desc_sw128(b + 32 * (u & 3) + 4096 * (u >> 2));
        const uint32_t d = tmem + (uint32_t)((u & amask) * spacing);
        if (elect_one()) mma_tf32(d, da, db, idesc, (i + u) >= nacc ? 1u : 0u);
      }
    }
    long long t1 = clock64();
    if (elect_one()) {
      umma_commit(&done);
      mbar_wait(&done, 0);
      long long t2 = clock64();
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = t2 - t0;
    }
  }
  // mode 0: the MMA loop runs in a lane-0 branch (tid == 0); mode 1: the whole warp runs the loop with warp-uniform
  // operands and only the tcgen05.mma itself is predicated by elect.sync
  if (mode < 2 && warp == 0 && (mode == 1 || tid == 0)) {
    const uint32_t a = smem_u32(tiles), b = a + 32 * 1024;
    const uint32_t idesc = kind_f16 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                                    : idesc_tf32(128, N);
    const int spacing = 512 / nacc, amask = nacc - 1;
    long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const int j = i & 3;
      const uint64_t da = swz ? make_desc_sw128(a + 32 * j) : make_desc(a + 2 * j * 2064, 2064, 128);
      const uint64_t db = swz ? make_desc_sw128(b + 32 * j) : make_desc(b + 2 * j * 2064, 2064, 128);
      const uint32_t d = tmem + (uint32_t)((i & amask) * spacing);
      bool go = true;
      if (mode == 1) go = elect_one();
      if (go) {
        if (kind_f16) {
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(i >= nacc ? 1u : 0u) : "memory");
        } else {
          mma_tf32(d, da, db, idesc, i >= nacc ? 1u : 0u);
        }
      }
    }
    long long t1 = clock64();
    if (tid == 0) {
      umma_commit(&done);
      mbar_wait(&done, 0);
      long long t2 = clock64();
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 2 * 148 * sizeof(long long));
  cudaFuncSetAttribute((const void*)rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int nmma = 256;
  for (int mode = 0; mode < 3; ++mode)
    for (int f16 = 0; f16 < (mode == 2 ? 1 : 2); ++f16)
      for (int N : {16, 48, 128, 256})
        for (int nacc : {1, 4}) {
          if (N * nacc > 512) continue;
          const int swz = 1, grid = 148;
          rate<<<grid, 128, 80 * 1024>>>(N, nacc, swz, nmma, f16, mode, out);
          rate<<<grid, 128, 80 * 1024>>>(N, nacc, swz, nmma, f16, mode, out);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2];
          cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
          printf("mode %d (%s) %s N=%3d nacc=%d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  (%s)\n", mode,
                 mode == 2 ? "warp-uniform, unrolled x8" : mode ? "warp-uniform + elect.sync" : "lane-0 branch", f16 ? "bf16 K16" : "tf32 K8 ", N, nacc,
                 (double)h[0] / nmma, (double)h[1] / nmma, cudaGetErrorString(e));
        }
  return 0;
}
