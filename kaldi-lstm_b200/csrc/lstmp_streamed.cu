// Weights-streamed mode: the elementwise kernels of the per-timestep path used when a layer's weight slices do
// NOT fit in shared memory across the SMs (e.g. BASELINE.json configs[4], 2048-cell / 1024-proj).  The recurrent
// and projection products then run as per-step tensor-core GEMMs that stream the weights from L2/HBM
// (lstmp_gemm_tc.cu), and these kernels do the fused gate / derivative arithmetic in between.  Same equations
// and the same activation record as the persistent kernels (LPS.h:261-325, 369-454).
#include "lstmp_common.cuh"
#include "lstmp_kernels.h"

namespace lstmp {

// one (stream, cell) per thread for frame t: gifo[t] holds x*W_x^T + bias + r(t-1)*W_r^T on entry
__global__ void __launch_bounds__(256) streamed_fwd_elem_kernel(float* __restrict__ gifo_t, const float* __restrict__ c_prev,
                                                                float* __restrict__ c_out, float* __restrict__ h_t,
                                                                float* __restrict__ m_t, const float* __restrict__ p_i,
                                                                const float* __restrict__ p_f,
                                                                const float* __restrict__ p_o, int S, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * C) return;
  const int s = idx / C, c = idx - s * C;
  float* gp = gifo_t + (size_t)s * 4 * C + c;
  const float cp = c_prev[idx];
  const float gi = sigmoidf_fast(gp[C] + cp * p_i[c]);        // :278,:284
  const float gf = sigmoidf_fast(gp[2 * C] + cp * p_f[c]);    // :281,:285
  const float gg = tanhf_fast(gp[0]);                         // :288
  float cc = gg * gi + cp * gf;                               // :291-294
  cc = fminf(fmaxf(cc, -kCellClip), kCellClip);               // :296-297
  const float h = tanhf_fast(cc);                             // :300
  const float go = sigmoidf_fast(gp[3 * C] + cc * p_o[c]);    // :303,:306
  gp[0] = gg; gp[C] = gi; gp[2 * C] = gf; gp[3 * C] = go;
  c_out[idx] = cc;
  h_t[idx] = h;
  m_t[idx] = h * go;                                          // :309
}

// derivative chain of frame t (LPS.h:411-440); dgifo_next / f_next / dc_next are NULL for the last frame
__global__ void __launch_bounds__(256) streamed_bwd_elem_kernel(const float* __restrict__ dm, const float* __restrict__ gifo_t,
                                                                const float* __restrict__ gifo_next,
                                                                const float* __restrict__ c_t, const float* __restrict__ c_prev,
                                                                const float* __restrict__ h_t,
                                                                const float* __restrict__ dgifo_next,
                                                                const float* __restrict__ dc_next, float* __restrict__ dgifo_t,
                                                                float* __restrict__ dc_t, const float* __restrict__ p_i,
                                                                const float* __restrict__ p_f,
                                                                const float* __restrict__ p_o, int S, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * C) return;
  const int s = idx / C, c = idx - s * C;
  const float* gp = gifo_t + (size_t)s * 4 * C + c;
  const float yg = gp[0], yi = gp[C], yf = gp[2 * C], yo = gp[3 * C];
  const float yh = h_t[idx], ycp = c_prev[idx];
  const float d_m = dm[idx];
  const float d_h = (d_m * yo) * (1.0f - yh * yh);
  const float d_o = (d_m * yh) * yo * (1.0f - yo);
  float d_c = d_h;
  if (dgifo_next) {
    const float* dn = dgifo_next + (size_t)s * 4 * C + c;
    d_c += dc_next[idx] * gifo_next[(size_t)s * 4 * C + 2 * C + c];   // :425
    d_c += dn[C] * p_i[c];                                            // :426
    d_c += dn[2 * C] * p_f[c];                                        // :427
  }
  d_c += d_o * p_o[c];                                                // :428
  float* dp = dgifo_t + (size_t)s * 4 * C + c;
  dp[0] = (d_c * yi) * (1.0f - yg * yg);
  dp[C] = (d_c * yg) * yi * (1.0f - yi);
  dp[2 * C] = (d_c * ycp) * yf * (1.0f - yf);
  dp[3 * C] = d_o;
  dc_t[idx] = d_c;
  (void)c_t;
}

// bias / peephole gradients (LPS.h:474-484): column sums over all T*S rows; one thread per column, fixed order.
__global__ void __launch_bounds__(128) streamed_small_grads_kernel(const float* __restrict__ dgifo, const float* __restrict__ cbuf,
                                                                   float* __restrict__ g_small, int rows, int S, int C) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= 7 * C) return;
  float a = 0.f;
  if (col < 4 * C) {
    for (int r = 0; r < rows; ++r) a += dgifo[(size_t)r * 4 * C + col];                       // bias
  } else {
    const int w = (col - 4 * C) / C, c = (col - 4 * C) - w * C;                               // 0: i, 1: f, 2: o
    const int gate = w + 1;                                                                   // DI, DF, DO columns
    const size_t coff = (w == 2) ? (size_t)S * C : 0;                                         // o uses c(t), i/f use c(t-1)
    for (int r = 0; r < rows; ++r) a += dgifo[(size_t)r * 4 * C + gate * C + c] * cbuf[coff + (size_t)r * C + c];
  }
  g_small[col] = a;
}

cudaError_t launch_streamed_fwd_elem(float* gifo_t, const float* c_prev, float* c_out, float* h_t, float* m_t,
                                     const float* p_i, const float* p_f, const float* p_o, int S, int C,
                                     cudaStream_t st) {
  streamed_fwd_elem_kernel<<<ceil_div(S * C, 256), 256, 0, st>>>(gifo_t, c_prev, c_out, h_t, m_t, p_i, p_f, p_o, S, C);
  return cudaGetLastError();
}
cudaError_t launch_streamed_bwd_elem(const float* dm, const float* gifo_t, const float* gifo_next, const float* c_t,
                                     const float* c_prev, const float* h_t, const float* dgifo_next,
                                     const float* dc_next, float* dgifo_t, float* dc_t, const float* p_i,
                                     const float* p_f, const float* p_o, int S, int C, cudaStream_t st) {
  streamed_bwd_elem_kernel<<<ceil_div(S * C, 256), 256, 0, st>>>(dm, gifo_t, gifo_next, c_t, c_prev, h_t, dgifo_next,
                                                                dc_next, dgifo_t, dc_t, p_i, p_f, p_o, S, C);
  return cudaGetLastError();
}
cudaError_t launch_streamed_small_grads(const float* dgifo, const float* cbuf, float* g_small, int rows, int S, int C,
                                        cudaStream_t st) {
  streamed_small_grads_kernel<<<ceil_div(7 * C, 128), 128, 0, st>>>(dgifo, cbuf, g_small, rows, S, C);
  return cudaGetLastError();
}

}  // namespace lstmp
