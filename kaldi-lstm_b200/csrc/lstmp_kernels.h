// Internal (non-ABI) interface between the host engine and the sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <vector>
#include <stddef.h>
#include <stdint.h>

namespace lstmp {

constexpr int kMaxGroupsHost = 8;
constexpr int kBarStride = 256;  // flags per group (>= CTAs per group), 1 KB apart

// Work decomposition of one persistent launch (shared by forward and backward).
//   grid = ngroups * ctas_per_group co-resident CTAs (cooperative launch, 1 CTA / SM).
//   Group g owns streams [g*Sg, (g+1)*Sg); inside a group CTA j owns
//     cells      [j*cpc, j*cpc+cpc)      (gate rows of W_gifo_r, cell-wise elementwise work)
//     r columns  [j*rpc, j*rpc+rpc)      (rows of W_r_m for the projection, forward only)
//     a `piece`-float slice of the group's flattened [Sg x R] d_r block (backward reduce-scatter).
struct Decomp {
  int ngroups, ctas_per_group, Sg, cpc, rpc, piece;
  int fwd_xcap, bwd_xcap;    // capacity (floats) of the shared-memory all-gather buffer
  int dbg;                   // timing experiments only (LSTMP_B200_DEBUG): 1 = no group barrier, 2 = no products,
                             // 4 = CTA 0 records clock64() stamps into dbg_stamps
  long long* dbg_stamps;     // host-mapped [2 + 2*1024]: count, pad, then (tag, clock) pairs
};

struct FwdParams {
  int I, C, R, S, T;
  Decomp d;
  // dynamic shared memory carve-up (offsets in floats)
  int off_wr, ldwr, off_wm, ldwm, off_xbuf, off_red, ldred, off_cprev, off_peep, off_plan;
  const float *w_gifo_r, *w_r_m, *p_i, *p_f, *p_o;
  float *gifo;   // [T*S x 4C]  in: x*W_x^T + bias (pre-activations); out: g,i,f,o activations
  float *cbuf;   // [(T+1)*S x C]  block 0 = c_0
  float *hbuf;   // [T*S x C]
  float *mbuf;   // [T*S x C]
  float *rbuf;   // [(T+1)*S x R]  block 0 = r_0
  float *out;    // [T*S x R], row stride ld_out
  long long ld_out;
  float *state_c;  // [S x C]
  float *state_r;  // [S x R]
  unsigned *bar;   // [ngroups][kBarStride] per-CTA epoch flags
  unsigned bar_base[kMaxGroupsHost];
};

struct BwdParams {
  int I, C, R, S, T;
  Decomp d;
  int off_wr, ldwr, off_wmt, ldwmt, off_xbuf, off_red, ldred, off_dgn, ldd, off_dcn, off_acc7, off_peep, off_plan;
  const float *w_gifo_r, *w_r_m, *p_i, *p_f, *p_o;
  const float *gifo, *cbuf, *hbuf;
  const float *out_diff;
  long long ld_od;
  float *dgifo;        // [T*S x 4C]
  float *dr;           // [T*S x R]
  float *scratch;      // [ngroups][ctas_per_group][Sg*R] partial d_r
  float *small_grads;  // [ngroups][7C]: bias(4C) | peephole_i | peephole_f | peephole_o
  unsigned *bar;
  unsigned bar_base[kMaxGroupsHost];
};

size_t fwd_smem_floats(int C, int R, Decomp& d, FwdParams* p, size_t limit_floats);
size_t bwd_smem_floats(int C, int R, Decomp& d, BwdParams* p, size_t limit_floats);
cudaError_t launch_fwd(const FwdParams& p, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_bwd(const BwdParams& p, size_t smem_bytes, cudaStream_t stream);
cudaError_t set_kernel_smem_limits(size_t fwd_bytes, size_t bwd_bytes);
int static_smem_reserve();
int fwd_barriers(int T);
int bwd_barriers(int T);

// ---- TMA-fed tcgen05 time loops, forward and backward (lstmp_recurrent_tma.cu) --------------------------------
// Activations cross the chip as bf16 hi/lo pairs written by their producer into global "tile image" arrays (per 64-k
// chunk the ready-made SWIZZLE_128B shared-memory image of [hi rows | lo rows]) and are pulled into a shared-memory
// ring with one cp.async.bulk per chunk; the CTA's weight slice is the stationary B operand.
struct FwdTmaParams {
  int I, C, R, S, T;
  int G, Sg, cpg;               // stream groups (own barrier each), streams per group, CTAs per group
  int nctas, cpc, rpc;          // CTA j of a group owns cells [j*cpc, +cpc) and r columns [j*rpc, +rpc)
  int n_g, n_p;                 // MMA N of the gate / projection products
  int nch_g, nch_p;             // 64-k chunks of the two contractions (K = R, K = C)
  int nslot, nprod, stagger;
  int gt;                       // 64-k tiles per ring slot = per bulk copy
  unsigned slot_bytes;
  unsigned chunk_g, chunk_p;    // bytes of one 64-k tile of the stationary weight slices
  unsigned off_bg, off_bp, off_ring, off_red, ldred, off_cprev, off_peep, off_bars;
  const float *w_gifo_r, *w_r_m, *p_i, *p_f, *p_o;
  float *gifo, *cbuf, *hbuf, *mbuf, *rbuf, *out;
  long long ld_out;
  float *state_c, *state_r;
  uint8_t *rhl, *mhl;           // [G][chunks][tile image]: bf16 hi/lo halves of the latest r / m (see store_hl)
  unsigned* bar;
  unsigned bar_base;
  int dbg;
  long long* dbg_stamps;
};
struct BwdTmaParams {
  int I, C, R, S, T;
  int G, Sg, cpg;               // stream groups, streams per group, CTAs per group (a multiple of kp)
  int nctas, kp;                // grid = nctas CTAs in clusters of kp; cluster b of a group owns d_r columns [b*rpb, +rpb)
  int cpc, rpb;
  int n_a, n_b;                 // MMA N of the d_r / d_m products
  int nch_a, nch_b;             // 64-k chunks: K = 4C (split over the kp ranks of a cluster), K = R
  int nslot, nprod, stagger;
  int gt;                       // 64-k tiles per ring slot = per bulk copy
  unsigned slot_bytes;
  unsigned chunk_a, chunk_b;
  unsigned off_ba, off_bb, off_ring, off_red, ldred, off_ext, off_peep, off_bars;
  const float *w_gifo_r, *w_r_m, *p_i, *p_f, *p_o;
  const float *gifo, *cbuf, *hbuf;
  const float* out_diff;
  long long ld_od;
  float *dgifo, *dr;
  float* g_small;               // [G][7C] bias(4C) | peephole_i | peephole_f | peephole_o (G == 1: the gradient arena)
  uint8_t *dghl, *drhl;         // [G][chunks][tile image]: bf16 hi/lo halves of the latest DGIFO / d_r
  unsigned* bar;
  unsigned bar_base;
  int dbg;
  long long* dbg_stamps;
};
bool fwd_tma_plan(int C, int R, int S, int G, int nctas, size_t smem_limit, FwdTmaParams* p, size_t* smem_bytes);
bool bwd_tma_plan(int C, int R, int S, int G, int nctas, int kp, size_t smem_limit, BwdTmaParams* p, size_t* smem_bytes);
cudaError_t tma_set_smem_limits(size_t fwd_bytes, size_t bwd_bytes);
int bwd_tma_max_ctas(int kp, size_t smem_bytes, int max_ctas);
cudaError_t launch_fwd_tma(const FwdTmaParams& p, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_bwd_tma(const BwdTmaParams& p, size_t smem_bytes, cudaStream_t stream);
inline int fwd_tma_barriers(int T) { return 2 * T; }

// C[M x N] = alpha * op(A) * op(B) + beta * C (+ bias[n] broadcast over rows), row-major with
// leading dimensions; op(A) is M x K, op(B) is K x N.  tA/tB: 0 = as stored, 1 = transposed.
cudaError_t launch_gemm(float* C, long long ldc, int M, int N, int K, float alpha, const float* A, long long lda,
                        int tA, const float* B, long long ldb, int tB, float beta, const float* bias,
                        cudaStream_t stream);

// tcgen05 GEMM on pre-split bf16 hi/lo tile images fed by bulk copies (lstmp_gemm_hl.cu); same contract as launch_gemm.
struct HlWorkspace {
  uint8_t *a = nullptr, *b = nullptr;   // tile images of op(A) [M x K] and op(B)^T [N x K]
  size_t a_cap = 0, b_cap = 0, a_bytes = 0;
  uint8_t* g = nullptr;                 // image arena of a group launch
  size_t g_cap = 0;
  struct GroupPlan {                    // split-K counts chosen for a group of shapes (key = n, then M, N, K of each)
    int key[13];
    int splits[4];
  };
  std::vector<GroupPlan> plans;
};
cudaError_t launch_gemm_hl(HlWorkspace* w, float* C, long long ldc, int M, int N, int K, float alpha, const float* A,
                           long long lda, int tA, const float* B, long long ldb, int tB, float beta, const float* bias,
                           cudaStream_t stream, bool* handled, float* ws, size_t ws_floats, int* nlaunch, bool reuse_a);
// Up to four independent products C_i = alpha_i * op(A_i) * op(B_i) + beta_i * C_i (+ bias_i) in three launches (all
// operand splits, all products, all split-K reduces).  tA / tB as in launch_gemm.
struct HlGemmDesc {
  float* C;
  long long ldc;
  int M, N, K;
  float alpha;
  const float* A;
  long long lda;
  int tA;
  const float* B;
  long long ldb;
  int tB;
  float beta;
  const float* bias;
};
cudaError_t launch_gemm_hl_group(HlWorkspace* w, const HlGemmDesc* d, int n, cudaStream_t stream, bool* handled,
                                 float* ws, size_t ws_floats, int* nlaunch);
void gemm_hl_free(HlWorkspace* w);

// corr = G + momentum*corr ; param -= lr*corr   over the flat arena
// clip > 0: corr is clamped element-wise to [-clip, clip] before the step (the standard/ component's gradient clip)
cudaError_t launch_update(float* params, float* corr, const float* grads, size_t n, float lr, float momentum,
                          float clip, cudaStream_t stream);
// G[bias | p_i | p_f | p_o] = sum over groups of small_grads
cudaError_t launch_small_grads(float* g_small /*7C contiguous in the arena*/, const float* small, int ngroups,
                               int n7c, cudaStream_t stream);
// zero state rows of flagged streams (flags as bit mask words, by value)
struct ResetMask {
  unsigned w[32];
};
cudaError_t launch_reset(float* state_c, int C, float* state_r, int R, int S, const ResetMask& m,
                         cudaStream_t stream);
// weights-streamed mode (lstmp_streamed.cu)
cudaError_t launch_streamed_fwd_elem(float* gifo_t, const float* c_prev, float* c_out, float* h_t, float* m_t,
                                     const float* p_i, const float* p_f, const float* p_o, int S, int C,
                                     cudaStream_t st);
cudaError_t launch_streamed_bwd_elem(const float* dm, const float* gifo_t, const float* gifo_next, const float* c_t,
                                     const float* c_prev, const float* h_t, const float* dgifo_next,
                                     const float* dc_next, float* dgifo_t, float* dc_t, const float* p_i,
                                     const float* p_f, const float* p_o, int S, int C, cudaStream_t st);
cudaError_t launch_streamed_small_grads(const float* dgifo, const float* cbuf, float* g_small, int rows, int S, int C,
                                        cudaStream_t st);
// strided 2-D copy helper for set/get of pitched caller matrices is done with cudaMemcpy2DAsync.

}  // namespace lstmp
