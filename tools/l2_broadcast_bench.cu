// Microbenchmark behind DESIGN.md's all-gather analysis: every CTA (one per SM) reads the SAME L2-resident block
// with ld.global.cg.v4 (the time loop's all-gather pattern), versus every CTA reading its own block, versus the
// same block walked from a per-CTA rotated start.  Prints bytes/clk/SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l2_broadcast_bench tools/l2_broadcast_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ float4 ld_cg(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// mode 0: same block, same order; 1: same block, start rotated per CTA; 2: private block per CTA
template <int U>
__global__ void __launch_bounds__(256, 1) bcast(const float4* __restrict__ buf, int n4, int mode, int reps,
                                                float* sink, long long* cycles) {
  const int tid = threadIdx.x;
  const float4* base = buf + (mode == 2 ? (size_t)blockIdx.x * n4 : 0);
  const int rot = (mode == 1) ? (int)(((long long)blockIdx.x * n4 / gridDim.x) & ~255) : 0;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    for (int i = 0; i < n4; i += 256 * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int idx = i + u * 256 + tid + rot;
        if (idx >= n4) idx -= n4;
        v[u] = ld_cg(base + idx);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

template <int U>
void run(const float4* buf, int n4, int nsm, float* sink, long long* cyc) {
  for (int mode = 0; mode < 3; ++mode) {
    const int reps = 20;
    bcast<U><<<nsm, 256>>>(buf, n4, mode, 2, sink, cyc);  // warm (L2 fill)
    bcast<U><<<nsm, 256>>>(buf, n4, mode, reps, sink, cyc);
    cudaDeviceSynchronize();
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    double bpc = (double)n4 * 16 * reps / (double)mx;
    printf("block %4d KB  in-flight %2d x16B/thread  mode %d (%s): %.1f B/clk/SM  (%lld cycles per pass)\n",
           n4 * 16 / 1024, U, mode, mode == 0 ? "same block, same order" : mode == 1 ? "same block, rotated start" : "private blocks",
           bpc, mx / reps);
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  const int sizes_kb[] = {128, 200};
  float* sink;
  long long* cyc;
  cudaMalloc(&sink, 4);
  cudaMalloc(&cyc, nsm * sizeof(long long));
  for (int kb : sizes_kb) {
    const int n4 = kb * 1024 / 16;
    float4* buf;
    cudaMalloc(&buf, (size_t)nsm * n4 * 16);
    cudaMemset(buf, 0, (size_t)nsm * n4 * 16);
    run<4>(buf, n4, nsm, sink, cyc);
    run<8>(buf, n4, nsm, sink, cyc);
    run<16>(buf, n4, nsm, sink, cyc);
    cudaFree(buf);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
