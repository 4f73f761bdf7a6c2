// Non-recurrent kernels of the engine: batched GEMMs outside the time loop (FP32 SIMT
// reference implementation; the tcgen05 path lives in lstmp_gemm_tc.cu), the fused
// momentum+SGD update, the bias/peephole gradient gather and the per-stream state reset.
#include "lstmp_common.cuh"
#include "lstmp_kernels.h"

namespace lstmp {

// -----------------------------------------------------------------------------------------
// C = alpha*op(A)*op(B) + beta*C (+bias).  Replaces the AddMatMat calls outside the time loop:
// LPS.h:246 (+:259 bias), :457, :468, :471, :486 (CuMatrixBase::AddMatMat, cu-matrix.cc:909-945).
// 128x64 block tile, 16-deep K slabs, 8x4 register tile per thread.
// -----------------------------------------------------------------------------------------
constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

__global__ void __launch_bounds__(GTHREADS) gemm_simt_kernel(float* __restrict__ Cm, long long ldc, int M, int N,
                                                             int K, float alpha, const float* __restrict__ A,
                                                             long long lda, int tA, const float* __restrict__ B,
                                                             long long ldb, int tB, float beta,
                                                             const float* __restrict__ bias) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int ty = tid / 16, tx = tid % 16;  // 16 x 16 thread grid: rows 8*ty.., cols 4*tx..
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += GBK) {
    // A tile: GBM x GBK
    for (int idx = tid; idx < GBM * GBK; idx += GTHREADS) {
      int m, k;
      if (tA) { k = idx / GBM; m = idx - k * GBM; } else { m = idx / GBK; k = idx - m * GBK; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = tA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      As[k][m] = v;
    }
    for (int idx = tid; idx < GBN * GBK; idx += GTHREADS) {
      int n, k;
      if (tB) { n = idx / GBK; k = idx - n * GBK; } else { k = idx / GBN; n = idx - k * GBN; }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = tB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][8 * ty]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][8 * ty + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][4 * tx]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + 8 * ty + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + 4 * tx + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (beta != 0.f) v += beta * Cm[(size_t)gm * ldc + gn];
      if (bias) v += bias[gn];
      Cm[(size_t)gm * ldc + gn] = v;
    }
  }
}

cudaError_t launch_gemm_simt(float* C, long long ldc, int M, int N, int K, float alpha, const float* A,
                             long long lda, int tA, const float* B, long long ldb, int tB, float beta,
                             const float* bias, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  dim3 grid(ceil_div(N, GBN), ceil_div(M, GBM)), block(GTHREADS);
  gemm_simt_kernel<<<grid, block, 0, stream>>>(C, ldc, M, N, K, alpha, A, lda, tA, B, ldb, tB, beta, bias);
  return cudaGetLastError();
}

// -----------------------------------------------------------------------------------------
// Update (LPS.h:465-487 momentum accumulation + :501-512 SGD step), fused over the flat arena
// [w_gifo_x | w_gifo_r | bias | peephole_i | peephole_f | peephole_o | w_r_m] (GetParams order,
// LPS.h:162-189):   corr = G + momentum * corr ;  param += -lr * corr
// G is the fresh (all-reduced when N>1) gradient written by backpropagate.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) update_kernel(float4* __restrict__ params, float4* __restrict__ corr,
                                                     const float4* __restrict__ grads, size_t n4, float lr,
                                                     float momentum, float clip) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 g = grads[i], c = corr[i], w = params[i];
    c.x = fmaf(momentum, c.x, g.x);
    c.y = fmaf(momentum, c.y, g.y);
    c.z = fmaf(momentum, c.z, g.z);
    c.w = fmaf(momentum, c.w, g.w);
    if (clip > 0.f) {  // standard/nnet/nnet-lstm-projected.h:469-493: element-wise clamp of the accumulated gradient
      c.x = fminf(fmaxf(c.x, -clip), clip);
      c.y = fminf(fmaxf(c.y, -clip), clip);
      c.z = fminf(fmaxf(c.z, -clip), clip);
      c.w = fminf(fmaxf(c.w, -clip), clip);
    }
    w.x = fmaf(-lr, c.x, w.x);
    w.y = fmaf(-lr, c.y, w.y);
    w.z = fmaf(-lr, c.z, w.z);
    w.w = fmaf(-lr, c.w, w.w);
    corr[i] = c;
    params[i] = w;
  }
}

cudaError_t launch_update(float* params, float* corr, const float* grads, size_t n, float lr, float momentum,
                          float clip, cudaStream_t stream) {
  size_t n4 = n / 4;  // arena length is a multiple of 4 (C % 4 == 0)
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  update_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<float4*>(params), reinterpret_cast<float4*>(corr),
                                            reinterpret_cast<const float4*>(grads), n4, lr, momentum, clip);
  return cudaGetLastError();
}

__global__ void small_grads_kernel(float* __restrict__ g_small, const float* __restrict__ small, int ngroups,
                                   int n7c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n7c) return;
  float a = 0.f;
  for (int g = 0; g < ngroups; ++g) a += small[(size_t)g * n7c + i];
  g_small[i] = a;
}
cudaError_t launch_small_grads(float* g_small, const float* small, int ngroups, int n7c, cudaStream_t stream) {
  small_grads_kernel<<<ceil_div(n7c, 256), 256, 0, stream>>>(g_small, small, ngroups, n7c);
  return cudaGetLastError();
}

// Reset (LPS.h:212-220): zero the carried state of every flagged stream.
__global__ void reset_kernel(float* __restrict__ state_c, int C, float* __restrict__ state_r, int R, int S,
                             const __grid_constant__ ResetMask m) {
  int s = blockIdx.x;
  if (s >= S || !((m.w[s >> 5] >> (s & 31)) & 1u)) return;
  for (int i = threadIdx.x; i < C; i += blockDim.x) state_c[(size_t)s * C + i] = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) state_r[(size_t)s * R + i] = 0.f;
}
cudaError_t launch_reset(float* state_c, int C, float* state_r, int R, int S, const ResetMask& m,
                         cudaStream_t stream) {
  reset_kernel<<<S, 256, 0, stream>>>(state_c, C, state_r, R, S, m);
  return cudaGetLastError();
}

}  // namespace lstmp
