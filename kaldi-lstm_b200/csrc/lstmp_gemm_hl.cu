// tcgen05 GEMM fed by bulk copies of PRE-SPLIT operands, sm_100a only.
//
//   C[M x N] = alpha * op(A) * op(B) + beta * C (+ bias[n])       fp32 in, fp32 out
//
// Same call sites as lstmp_gemm_tc.cu (the contractions outside the time loop, LPS.h:246, :457, :468, :471, :486, and
// the three GEMMs of the output tail) and the same FP32-faithful idea as the TMA-fed time loops: every fp32 operand x is
// split once into two bf16 pieces, hi = bf16(x) and lo = bf16(x - hi), and the product is accumulated in FP32 in TMEM as
// hi*hi + hi*lo + lo*hi + lo*lo with tcgen05.mma.kind::f16 (K = 16 per instruction: half the instructions of the
// 3xTF32 kernel for the same K).  What changes is WHERE the split happens:
//
//   1. split_hl_kernel (one elementwise pass per operand, HBM-bound) writes the operand as "tile images": for every
//      [128 rows x 64 k] tile the ready-made SWIZZLE_128B K-major shared-memory image of its hi half and of its lo
//      half (16 KB each), zero-padded to whole tiles; an operand stored with the contraction index as the ROW index
//      (m/n-contiguous, e.g. DGIFO^T for the weight gradients) is transposed on the way.
//   2. gemm_hl_kernel (persistent, 320 threads): one thread issues two 32 KB cp.async.bulk copies per 64-deep K block
//      (A_hi | A_lo, B_hi | B_lo) straight into the UMMA ring (mbarrier expect_tx), one warp issues 16 MMAs per block
//      into one of two TMEM accumulators, eight warps drain the other one.  No loader warps, no register pass, no
//      generic-proxy stores into operand tiles, no bounds logic on the operand side.
//   3. launch_gemm_hl_group: up to four independent products -- the contractions after a layer's backward time loop --
//      as ONE split launch (split_multi_kernel), ONE persistent product launch and ONE split-K reduce launch.
//
// (lstmp_gemm_tc.cu moves 32 KB in, 32 KB back out and 64 KB of hi/lo tiles through shared memory per 32-deep K block
// with eight loader warps and reaches ~1.9 k cycles per block against 0.78 k of MMA time)
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "lstmp_common.cuh"
#include "lstmp_gemm_plan.h"
#include "lstmp_kernels.h"
#include "lstmp_tc.cuh"

namespace lstmp {
namespace hl {
constexpr int BM = 128, BN = 128, BK = 64;           // tile: 128 x 128 outputs, 64 k per stage
constexpr uint32_t TILE = 128 * 128;                 // bytes of one [128 rows x 64 k] bf16 image (hi or lo)
constexpr int NSTAGE = 3;                            // stages of A_hi | A_lo | B_hi | B_lo = 64 KB
constexpr uint32_t STAGE = 4 * TILE;
constexpr int THREADS = 320;                         // warp 0 producer, warp 1 MMA issuer + TMEM, warps 2-9 epilogue
constexpr int PLD = 20;                              // row stride (floats) of an epilogue warp's [32 x 16] staging patch
constexpr uint32_t PATCH = 32 * PLD * 4;
constexpr int NEPI = 8;                              // epilogue warps: two per TMEM lane quadrant (64 columns each)
constexpr uint32_t SMEM = NSTAGE * STAGE + NEPI * PATCH + 1024 + 256;

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Operand [R x K] (R = M or N) -> images img[(rt * nkt + kt) * 2 + {hi, lo}][TILE], rt = r / 128, kt = k / 64.
// TRANSPOSED == false: element (r, k) at src[r * ld + k]; true: at src[k * ld + r].
// One thread per 16-byte unit (8 consecutive k of one row); whole tiles are written (zeros outside R x K).
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256) split_hl_kernel(uint8_t* __restrict__ img, const float* __restrict__ src,
                                                       long long ld, int R, int K, int nrt, int nkt) {
  const long long units = (long long)nrt * nkt * 128 * 8;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units;
       u += (long long)gridDim.x * blockDim.x) {
    int rl, kc, kt, rt;
    if (!TRANSPOSED) {  // consecutive threads: consecutive k units of a row (coalesced reads along k)
      kc = (int)(u & 7);
      long long t = u >> 3;
      kt = (int)(t % nkt);
      t /= nkt;
      rl = (int)(t & 127);
      rt = (int)(t >> 7);
    } else {            // consecutive threads: consecutive rows (= consecutive source columns: coalesced reads)
      rl = (int)(u & 127);
      long long t = u >> 7;
      rt = (int)(t % nrt);
      t /= nrt;
      kc = (int)(t & 7);
      kt = (int)(t >> 3);
    }
    const int r = rt * 128 + rl, k0 = kt * BK + kc * 8;
    float v[8];
    if (!TRANSPOSED) {
      if (r < R && k0 + 7 < K && ((ld & 3) == 0)) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + (long long)r * ld + k0));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + (long long)r * ld + k0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (r < R && k0 + q < K) ? __ldg(src + (long long)r * ld + k0 + q) : 0.f;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (r < R && k0 + q < K) ? __ldg(src + (long long)(k0 + q) * ld + r) : 0.f;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    uint8_t* t = img + ((size_t)rt * nkt + kt) * 2 * TILE + tc::sw128_off(rl, kc);
    *reinterpret_cast<uint4*>(t) = hi;
    *reinterpret_cast<uint4*>(t + TILE) = lo;
  }
}

// Both operands of a GEMM in ONE launch (the direct kernel for each): blocks [0, nblk_a) work on A, the rest on B.
struct SplitJob {
  uint8_t* img;
  const float* src;
  long long ld;
  int R, K, nrt, nkt, transposed;
};
template <bool TRANSPOSED>
__device__ __forceinline__ void split_units(const SplitJob& j, long long first, long long stride) {
  const long long units = (long long)j.nrt * j.nkt * 128 * 8;
  for (long long u = first; u < units; u += stride) {
    int rl, kc, kt, rt;
    if (!TRANSPOSED) {
      kc = (int)(u & 7);
      long long t = u >> 3;
      kt = (int)(t % j.nkt);
      t /= j.nkt;
      rl = (int)(t & 127);
      rt = (int)(t >> 7);
    } else {
      rl = (int)(u & 127);
      long long t = u >> 7;
      rt = (int)(t % j.nrt);
      t /= j.nrt;
      kc = (int)(t & 7);
      kt = (int)(t >> 3);
    }
    const int r = rt * 128 + rl, k0 = kt * BK + kc * 8;
    float v[8];
    if (!TRANSPOSED) {
      if (r < j.R && k0 + 7 < j.K && ((j.ld & 3) == 0)) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(j.src + (long long)r * j.ld + k0));
        const float4 b = __ldg(reinterpret_cast<const float4*>(j.src + (long long)r * j.ld + k0 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (r < j.R && k0 + q < j.K) ? __ldg(j.src + (long long)r * j.ld + k0 + q) : 0.f;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (r < j.R && k0 + q < j.K) ? __ldg(j.src + (long long)(k0 + q) * j.ld + r) : 0.f;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    uint8_t* t = j.img + ((size_t)rt * j.nkt + kt) * 2 * TILE + tc::sw128_off(rl, kc);
    *reinterpret_cast<uint4*>(t) = hi;
    *reinterpret_cast<uint4*>(t + TILE) = lo;
  }
}
__global__ void __launch_bounds__(256) split_pair_kernel(const SplitJob ja, const SplitJob jb, int nblk_a) {
  const bool is_a = (int)blockIdx.x < nblk_a;
  const SplitJob& j = is_a ? ja : jb;
  const long long nb = is_a ? nblk_a : (long long)gridDim.x - nblk_a;
  const long long first = ((long long)(is_a ? blockIdx.x : blockIdx.x - nblk_a)) * blockDim.x + threadIdx.x;
  if (j.transposed) split_units<true>(j, first, nb * blockDim.x);
  else split_units<false>(j, first, nb * blockDim.x);
}

// TRANSPOSED operand (element (r, k) at src[k * ld + r]) through a shared-memory tile: one CTA per [128 r x 64 k] image
// tile reads 64 source rows of 512 contiguous bytes (coalesced along r) and writes, per image row, the 8 hi units and
// the 8 lo units = two whole 128-byte lines.  (The direct version above read coalesced but wrote 16-byte pieces to 32
// different lines per warp: 11.4 us for DGIFO^T [3200 x 1280] against 6.5 us of HBM time.)
__device__ __forceinline__ void split_tile_transposed(uint8_t* __restrict__ img, const float* __restrict__ src,
                                                      long long ld, int R, int K, int nrt, int nkt, int tile_id,
                                                      float (*tile)[129]) {
  const int rt = tile_id % nrt, kt = tile_id / nrt;
  const int r0 = rt * 128, k0 = kt * BK;
  const int tid = threadIdx.x;
  for (int i = tid; i < BK * 128; i += 256) {
    const int kk = i >> 7, rl = i & 127;
    const int r = r0 + rl, k = k0 + kk;
    tile[kk][rl] = (r < R && k < K) ? __ldg(src + (long long)k * ld + r) : 0.f;
  }
  __syncthreads();
  uint8_t* t = img + ((size_t)rt * nkt + kt) * 2 * TILE;
  for (int u = tid; u < 128 * 8; u += 256) {
    const int kc = u & 7, rl = u >> 3;  // consecutive threads: the 8 units of one image row (one 128-byte line)
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = tile[kc * 8 + q][rl];
    uint4 hi, lo;
    split8(v, hi, lo);
    const uint32_t off = tc::sw128_off(rl, kc);
    *reinterpret_cast<uint4*>(t + off) = hi;
    *reinterpret_cast<uint4*>(t + TILE + off) = lo;
  }
}
__global__ void __launch_bounds__(256) split_hl_transposed_tiled_kernel(uint8_t* __restrict__ img,
                                                                        const float* __restrict__ src, long long ld,
                                                                        int R, int K, int nrt, int nkt) {
  __shared__ float tile[BK][129];
  split_tile_transposed(img, src, ld, R, K, nrt, nkt, blockIdx.x, tile);
}

// Every operand of a GROUP of GEMMs in ONE launch: blocks [blk0[i], blk0[i + 1]) work on job i; SplitJob::transposed
// = 0 direct, 1 direct transposed, 2 transposed through the shared-memory tile (one [128 x 64] image tile at a time).
constexpr int MAX_SPLIT_JOBS = 8;
struct SplitJobs {
  SplitJob j[MAX_SPLIT_JOBS];
  int blk0[MAX_SPLIT_JOBS + 1];
  int n;
};
__global__ void __launch_bounds__(256) split_multi_kernel(const __grid_constant__ SplitJobs g) {
  __shared__ float tile[BK][129];
  int i = 0;
#pragma unroll
  for (int q = 1; q < MAX_SPLIT_JOBS; ++q)
    if (q < g.n && (int)blockIdx.x >= g.blk0[q]) i = q;
  const SplitJob& j = g.j[i];
  const int lb = (int)blockIdx.x - g.blk0[i], nb = g.blk0[i + 1] - g.blk0[i];
  if (j.transposed == 2) {
    const int ntiles = j.nrt * j.nkt;
    for (int t = lb; t < ntiles; t += nb) {   // (block-uniform trip count)
      split_tile_transposed(j.img, j.src, j.ld, j.R, j.K, j.nrt, j.nkt, t, tile);
      __syncthreads();                          // the tile is refilled in the next round
    }
  } else if (j.transposed == 1) {
    split_units<true>(j, (long long)lb * blockDim.x + threadIdx.x, (long long)nb * blockDim.x);
  } else {
    split_units<false>(j, (long long)lb * blockDim.x + threadIdx.x, (long long)nb * blockDim.x);
  }
}

// Split-K partial sums of up to MAX_GROUP products -> C, in a fixed order (deterministic), one launch.
constexpr int MAX_GROUP = 4;
static_assert(BM == hlplan::kBM && BN == hlplan::kBN && BK == hlplan::kBK && MAX_GROUP == hlplan::kMaxGroup,
              "lstmp_gemm_plan.h plans for this kernel's tile");
struct ReduceJob {
  float* C;
  const float* ws;
  const float* bias;
  long long ldc;
  int M, N, splits;
  float alpha, beta;
};
struct ReduceJobs {
  ReduceJob j[MAX_GROUP];
  int blk0[MAX_GROUP + 1];
  int n;
};
__global__ void __launch_bounds__(256) splitk_reduce_multi_kernel(const __grid_constant__ ReduceJobs g) {
  int i = 0;
#pragma unroll
  for (int q = 1; q < MAX_GROUP; ++q)
    if (q < g.n && (int)blockIdx.x >= g.blk0[q]) i = q;
  const ReduceJob& r = g.j[i];
  const int lb = (int)blockIdx.x - g.blk0[i], nb = g.blk0[i + 1] - g.blk0[i];
  const int n4 = r.N >> 2;
  const long long total = (long long)r.M * n4;
  for (long long idx = (long long)lb * blockDim.x + threadIdx.x; idx < total; idx += (long long)nb * blockDim.x) {
    const int m = (int)(idx / n4), n = (int)(idx - (long long)m * n4) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < r.splits; ++z) {
      const float4 v = *reinterpret_cast<const float4*>(r.ws + ((size_t)z * r.M + m) * r.N + n);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    float* c = r.C + (size_t)m * r.ldc + n;
    float o[4] = {r.alpha * a.x, r.alpha * a.y, r.alpha * a.z, r.alpha * a.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (r.beta != 0.f) o[q] += r.beta * c[q];
      if (r.bias) o[q] += r.bias[n + q];
      c[q] = o[q];
    }
  }
}

__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}

// One product of a group: C[M x N] = alpha * op(A) * op(B) + beta * C (+ bias); work items [item0, item0 + ntm * ntn *
// splits) of the launch, item = (m tile, n tile, K split) with m fastest.
struct HlProblem {
  float* C;
  const float* bias;
  const uint8_t* a_img;   // tile images of op(A) [M x K] and op(B)^T [N x K]
  const uint8_t* b_img;
  float* split_ws;        // splits > 1: raw partial sums [splits][M x N]; splitk_reduce applies alpha / beta / bias
  long long ldc;
  int M, N, nkt, ntm, ntn, splits, kt_per_split, item0;
  float alpha, beta;
};
struct HlGroup {
  HlProblem p[MAX_GROUP];
  int n, nitems;
};
struct HlItem {
  int pi, mt, nt, z;
};
__device__ __forceinline__ HlItem decode_item(const HlGroup& g, int w) {
  HlItem it;
  it.pi = 0;
#pragma unroll
  for (int q = 1; q < MAX_GROUP; ++q)
    if (q < g.n && w >= g.p[q].item0) it.pi = q;
  const HlProblem& P = g.p[it.pi];
  const int l = w - P.item0;
  it.mt = l % P.ntm;
  it.nt = (l / P.ntm) % P.ntn;
  it.z = l / (P.ntm * P.ntn);
  return it;
}

// PERSISTENT: grid = min(work items, #SMs) CTAs over the work items of UP TO FOUR independent products (the GEMMs that
// follow a layer's backward time loop run as one launch); within a product m is fastest (the CTAs that run at the same
// time share the B tile in L2), item w of CTA c = c + i * gridDim.x.  The accumulator is double-buffered in TMEM
// (2 x 128 columns): the eight epilogue warps drain tile i (tcgen05.ld -> alpha/beta/bias -> 64 contiguous bytes
// per 4 lanes) while the producer / MMA warps are already in the main loop of tile i+1; the shared-memory ring
// runs continuously across tiles and products.
__global__ void __launch_bounds__(THREADS, 1) gemm_hl_kernel(const __grid_constant__ HlGroup g) {
  extern __shared__ __align__(128) uint8_t smem_raw_hl[];
  uint8_t* tiles = smem_raw_hl + ((1024u - (smem_u32(smem_raw_hl) & 1023u)) & 1023u);
  float* patches = reinterpret_cast<float*>(tiles + NSTAGE * STAGE);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + NSTAGE * STAGE + NEPI * PATCH);
  uint64_t* empty = full + NSTAGE;
  uint64_t* acc_full = empty + NSTAGE;   // [2] accumulator buffer complete (MMA warp -> epilogue)
  uint64_t* acc_empty = acc_full + 2;    // [2] accumulator buffer drained (NEPI epilogue warps -> MMA warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int nitems = g.nitems;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], NEPI);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "n"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tiles_s = smem_u32(tiles);

  if (warp == 0) {
    // ------------------------------ producer: 2 bulk copies of 32 KB per K block -------------
    if (lane == 0) {
      uint32_t kbc = 0;  // K blocks issued so far (ring position, continuous across work items)
      for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
        const HlItem wi = decode_item(g, w);
        const HlProblem& P = g.p[wi.pi];
        const int nkt = P.nkt, kt0 = wi.z * P.kt_per_split, nk = min(P.kt_per_split, nkt - kt0);
        const uint8_t* ap = P.a_img + ((size_t)wi.mt * nkt + kt0) * 2 * TILE;   // A_hi | A_lo of (mt, kt) are adjacent
        const uint8_t* bp = P.b_img + ((size_t)wi.nt * nkt + kt0) * 2 * TILE;
        for (int kb = 0; kb < nk; ++kb, ++kbc) {
          const uint32_t s = kbc % NSTAGE, u = kbc / NSTAGE;
          if (u > 0) mbar_wait(&empty[s], (u - 1) & 1);
          mbar_arrive_expect_tx(&full[s], STAGE);
          const uint32_t dst = tiles_s + s * STAGE;
          bulk_g2s(dst, ap + (size_t)kb * 2 * TILE, 2 * TILE, &full[s]);
          bulk_g2s(dst + 2 * TILE, bp + (size_t)kb * 2 * TILE, 2 * TILE, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer -----------------------------------------------
    // kind::f16, bf16 x bf16 -> fp32, both operands K-major: D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    uint32_t kbc = 0, it = 0;
    for (int w = blockIdx.x; w < nitems; w += gridDim.x, ++it) {
      const HlItem wi = decode_item(g, w);
      const HlProblem& P = g.p[wi.pi];
      const int kt0 = wi.z * P.kt_per_split, nk = min(P.kt_per_split, P.nkt - kt0);
      const uint32_t buf = it & 1, ub = it >> 1;
      if (ub > 0) {  // the epilogue has drained this accumulator buffer (its previous tile)
        mbar_wait(&acc_empty[buf], (ub - 1) & 1);
        tc::tc_fence_after();
      }
      const uint32_t tmem_d = tmem_base + buf * BN;
      for (int kb = 0; kb < nk; ++kb, ++kbc) {
        const uint32_t s = kbc % NSTAGE, u = kbc / NSTAGE;
        mbar_wait(&full[s], u & 1);
        tc::tc_fence_after();
        {
          const uint32_t a_hi = tiles_s + s * STAGE, a_lo = a_hi + TILE, b_hi = a_hi + 2 * TILE, b_lo = b_hi + TILE;
#pragma unroll
          for (int j = 0; j < BK / 16; ++j) {
            const uint64_t dah = tc::make_desc_sw128(a_hi + 32 * j), dal = tc::make_desc_sw128(a_lo + 32 * j);
            const uint64_t dbh = tc::make_desc_sw128(b_hi + 32 * j), dbl = tc::make_desc_sw128(b_lo + 32 * j);
            if (tc::elect_one()) mma_bf16(tmem_d, dal, dbl, idesc, (kb | j) ? 1u : 0u);  // small terms first
            if (tc::elect_one()) mma_bf16(tmem_d, dal, dbh, idesc, 1u);
            if (tc::elect_one()) mma_bf16(tmem_d, dah, dbl, idesc, 1u);
            if (tc::elect_one()) mma_bf16(tmem_d, dah, dbh, idesc, 1u);
          }
          if (tc::elect_one()) {
            tc::umma_commit(&empty[s]);
            if (kb == nk - 1) tc::umma_commit(&acc_full[buf]);
          }
        }
        __syncwarp();
      }
    }
    tc::tc_fence_before();
  } else {
    // ------------------------------ epilogue (warps 2-9: TMEM lane quadrants 2, 3, 0, 1, 2, 3, 0, 1) -----
    const int quad = warp & 3;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < nitems; w += gridDim.x, ++it) {
      const HlItem wi = decode_item(g, w);
      const HlProblem& P = g.p[wi.pi];
      const int M = P.M, N = P.N;
      const int m0 = wi.mt * BM, n0 = wi.nt * BN;
      float* Cm = P.C;
      long long ldc = P.ldc;
      float alpha = P.alpha, beta = P.beta;
      const float* bias = P.bias;
      if (P.splits > 1) {  // split-K: raw partial sums to split_ws[z][M x N]; the reduce applies alpha / beta / bias
        Cm = P.split_ws + (size_t)wi.z * M * N;
        ldc = N;
        alpha = 1.f;
        beta = 0.f;
        bias = nullptr;
      }
      const uint32_t buf = it & 1, ub = it >> 1;
      mbar_wait(&acc_full[buf], ub & 1);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(quad * 32) << 16);
      float* patch = patches + (size_t)(warp - 2) * (PATCH / 4);
      // 16 columns at a time: TMEM lane (= row) -> this warp's staging patch -> global memory with 4 lanes per row
      // (64 contiguous bytes), 8 rows per store instruction.  Two warps share a lane quadrant (column halves).
      const int half = (warp - 2) >> 2;
      const int rsub = lane >> 2, c4 = (lane & 3) * 4;
#pragma unroll 1
      for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 16) {
        float v[16];
        tc::tmem_ld16(taddr + (uint32_t)c, v);
        float* prow = patch + lane * PLD;
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(prow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int gn = n0 + c + c4;
        const bool vec_ok = ((ldc & 3) == 0) && gn + 3 < N;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && vec_ok) bb = __ldg(reinterpret_cast<const float4*>(bias + gn));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rl = rsub + 8 * i;
          const int gm = m0 + quad * 32 + rl;
          if (gm < M) {
            const float4 a4 = *reinterpret_cast<const float4*>(patch + rl * PLD + c4);
            float* crow = Cm + (size_t)gm * ldc;
            if (vec_ok) {
              float4 o = make_float4(alpha * a4.x, alpha * a4.y, alpha * a4.z, alpha * a4.w);
              if (beta != 0.f) {
                const float4 cc = *reinterpret_cast<const float4*>(crow + gn);
                o.x += beta * cc.x; o.y += beta * cc.y; o.z += beta * cc.z; o.w += beta * cc.w;
              }
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
              *reinterpret_cast<float4*>(crow + gn) = o;
            } else {
              const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (gn + q < N) {
                  float o = alpha * av[q];
                  if (beta != 0.f) o += beta * crow[gn + q];
                  if (bias) o += bias[gn + q];
                  crow[gn + q] = o;
                }
              }
            }
          }
        }
        __syncwarp();  // the patch is rewritten in the next iteration
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);  // this warp's quadrant of the buffer is free again
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}
}  // namespace hl

cudaError_t launch_splitk_reduce(float* C, long long ldc, int M, int N, float alpha, float beta, const float* bias,
                                 const float* ws, int splits, cudaStream_t stream);  // lstmp_gemm_tc.cu

size_t gemm_hl_image_bytes(int rows, int K) {
  return (size_t)((rows + 127) / 128) * ((K + hl::BK - 1) / hl::BK) * 2 * hl::TILE;
}

// src: the operand as stored; rows x K is its logical [R x K] shape; transposed: element (r, k) at src[k * ld + r].
cudaError_t gemm_hl_split(uint8_t* img, const float* src, long long ld, int rows, int K, bool transposed,
                          cudaStream_t stream) {
  const int nrt = (rows + 127) / 128, nkt = (K + hl::BK - 1) / hl::BK;
  const long long units = (long long)nrt * nkt * 128 * 8;
  int blocks = (int)((units + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  // transposed sources: the tiled kernel writes whole 128-byte lines (3.4-4.6 TB/s on the 34-42 MB operands of the
  // tail) but needs a few hundred tiles to fill the GPU; small operands take the direct kernel (5 vs 8 us at 40 tiles)
  if (transposed && nrt * nkt >= 2 * 148)
    hl::split_hl_transposed_tiled_kernel<<<nrt * nkt, 256, 0, stream>>>(img, src, ld, rows, K, nrt, nkt);
  else if (transposed) hl::split_hl_kernel<true><<<blocks, 256, 0, stream>>>(img, src, ld, rows, K, nrt, nkt);
  else hl::split_hl_kernel<false><<<blocks, 256, 0, stream>>>(img, src, ld, rows, K, nrt, nkt);
  return cudaGetLastError();
}

static cudaError_t gemm_hl_prepare(int* nsm) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!attr_set[dev & 63]) {
    e = cudaFuncSetAttribute((const void*)hl::gemm_hl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hl::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  *nsm = 148;
  return cudaDeviceGetAttribute(nsm, cudaDevAttrMultiProcessorCount, dev);
}

static hl::HlProblem make_problem(float* C, long long ldc, int M, int N, int K, float alpha, const uint8_t* a_img,
                                  const uint8_t* b_img, float beta, const float* bias, int splits, float* split_ws) {
  hl::HlProblem P;
  P.C = C; P.bias = bias; P.a_img = a_img; P.b_img = b_img; P.ldc = ldc;
  P.M = M; P.N = N; P.alpha = alpha; P.beta = beta;
  const hlplan::Tiling t = hlplan::tiling(M, N, K, splits);
  P.nkt = t.nkt;
  P.ntm = t.ntm;
  P.ntn = t.ntn;
  P.kt_per_split = t.kt_per_split;
  P.splits = t.splits;
  P.split_ws = P.splits > 1 ? split_ws : nullptr;
  P.item0 = 0;
  return P;
}

cudaError_t gemm_hl_run(float* C, long long ldc, int M, int N, int K, float alpha, const uint8_t* a_img,
                        const uint8_t* b_img, float beta, const float* bias, cudaStream_t stream, float* ws,
                        size_t ws_floats, int* nlaunch) {
  *nlaunch = 1;
  int nsm = 148;
  cudaError_t e = gemm_hl_prepare(&nsm);
  if (e != cudaSuccess) return e;
  const int nkt = (K + hl::BK - 1) / hl::BK;
  const int tiles = ((M + hl::BM - 1) / hl::BM) * ((N + hl::BN - 1) / hl::BN);
  int splits = 1;
  // split-K when the output tiles alone cannot fill the SMs (in_diff: 40 tiles, G(w_r_m): 28 tiles)
  if (ws && tiles < 100 && nkt >= 4 && (N & 3) == 0) {
    splits = nsm / tiles;
    if (splits > 8) splits = 8;
    if (splits > nkt / 2) splits = nkt / 2;
    while (splits > 1 && (size_t)splits * M * N > ws_floats) --splits;
  }
  hl::HlGroup g;
  memset(&g, 0, sizeof(g));
  g.p[0] = make_problem(C, ldc, M, N, K, alpha, a_img, b_img, beta, bias, splits, ws);
  g.n = 1;
  g.nitems = g.p[0].ntm * g.p[0].ntn * g.p[0].splits;
  dim3 grid(g.nitems < nsm ? g.nitems : nsm), block(hl::THREADS);
  hl::gemm_hl_kernel<<<grid, block, hl::SMEM, stream>>>(g);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (g.p[0].splits > 1) {
    *nlaunch = 2;
    return launch_splitk_reduce(C, ldc, M, N, alpha, beta, bias, ws, g.p[0].splits, stream);
  }
  return cudaSuccess;
}

static hl::SplitJob make_job(uint8_t* img, const float* src, long long ld, int rows, int K, bool transposed, int* blocks,
                             bool* tiled) {
  hl::SplitJob j;
  j.img = img; j.src = src; j.ld = ld; j.R = rows; j.K = K;
  j.nrt = (rows + 127) / 128;
  j.nkt = (K + hl::BK - 1) / hl::BK;
  j.transposed = transposed ? 1 : 0;
  const long long units = (long long)j.nrt * j.nkt * 128 * 8;
  int b = (int)((units + 255) / 256);
  if (b > 148 * 8) b = 148 * 8;
  *blocks = b;
  *tiled = transposed && j.nrt * j.nkt >= 2 * 148;
  return j;
}

// One call = split both operands (unless the caller says op(A)'s image of the previous call is still valid), run the
// GEMM.  The image buffers grow on demand and belong to the caller's handle (one stream at a time per handle).
cudaError_t launch_gemm_hl(HlWorkspace* w, float* C, long long ldc, int M, int N, int K, float alpha, const float* A,
                           long long lda, int tA, const float* B, long long ldb, int tB, float beta, const float* bias,
                           cudaStream_t stream, bool* handled, float* ws, size_t ws_floats, int* nlaunch, bool reuse_a) {
  *handled = false;
  *nlaunch = 0;
  if (!w || M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(C) || (bias && !al16(bias)) || !al16(A) || !al16(B)) return cudaSuccess;
  const size_t na = gemm_hl_image_bytes(M, K), nb = gemm_hl_image_bytes(N, K);
  auto grow = [&](uint8_t** p, size_t* cap, size_t need) -> cudaError_t {
    if (need <= *cap) return cudaSuccess;
    if (*p) {
      cudaError_t e = cudaFree(*p);  // (waits for the device: earlier kernels may still read the old buffer)
      if (e != cudaSuccess) return e;
      *p = nullptr;
      *cap = 0;
    }
    cudaError_t e = cudaMalloc((void**)p, need);
    if (e == cudaSuccess) *cap = need;
    return e;
  };
  const bool need_a = !(reuse_a && w->a && w->a_bytes == na);
  cudaError_t e = cudaSuccess;
  if (need_a && (e = grow(&w->a, &w->a_cap, na)) != cudaSuccess) return e;
  if ((e = grow(&w->b, &w->b_cap, nb)) != cudaSuccess) return e;
  int ba = 0, bb = 0;
  bool ta = false, tb = false;
  const hl::SplitJob ja = make_job(w->a, A, lda, M, K, tA != 0, &ba, &ta);
  const hl::SplitJob jb = make_job(w->b, B, ldb, N, K, tB == 0, &bb, &tb);
  if (need_a && !ta && !tb) {
    // both operands in one launch
    hl::split_pair_kernel<<<ba + bb, 256, 0, stream>>>(ja, jb, ba);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    w->a_bytes = na;
    ++*nlaunch;
  } else {
    if (need_a) {
      if ((e = gemm_hl_split(w->a, A, lda, M, K, tA != 0, stream)) != cudaSuccess) return e;
      w->a_bytes = na;
      ++*nlaunch;
    }
    if ((e = gemm_hl_split(w->b, B, ldb, N, K, tB == 0, stream)) != cudaSuccess) return e;
    ++*nlaunch;
  }
  int nl = 0;
  e = gemm_hl_run(C, ldc, M, N, K, alpha, w->a, w->b, beta, bias, stream, ws, ws_floats, &nl);
  *nlaunch += nl;
  *handled = true;
  return e;
}

// ---- a GROUP of products in three launches: every operand split (one launch), every product (one persistent launch),
// every split-K reduce (one launch).  The GEMMs that follow a layer's backward time loop -- in_diff, G(w_gifo_x),
// G(w_gifo_r), G(w_r_m) -- are independent of each other; as separate calls they were 9-10 launches of 4-20 us.
cudaError_t launch_gemm_hl_group(HlWorkspace* w, const HlGemmDesc* d, int n, cudaStream_t stream, bool* handled,
                                 float* ws, size_t ws_floats, int* nlaunch) {
  *handled = false;
  *nlaunch = 0;
  if (!w || n < 1 || n > hl::MAX_GROUP) return cudaSuccess;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  for (int i = 0; i < n; ++i) {
    if (d[i].M <= 0 || d[i].N <= 0 || d[i].K <= 0) return cudaSuccess;
    if (!al16(d[i].C) || !al16(d[i].A) || !al16(d[i].B) || (d[i].bias && !al16(d[i].bias))) return cudaSuccess;
  }
  int nsm = 148;
  cudaError_t e = gemm_hl_prepare(&nsm);
  if (e != cudaSuccess) return e;

  // ---- operands (an operand used by two products -- DGIFO^T -- is split once)
  struct Operand {
    const float* src;
    long long ld;
    int R, K;
    bool transposed;
    size_t off;
  };
  Operand ops[2 * hl::MAX_GROUP];
  int nops = 0, a_of[hl::MAX_GROUP], b_of[hl::MAX_GROUP];
  size_t bytes = 0;
  auto operand = [&](const float* src, long long ld, int R, int K, bool tr) {
    for (int i = 0; i < nops; ++i)
      if (ops[i].src == src && ops[i].ld == ld && ops[i].R == R && ops[i].K == K && ops[i].transposed == tr) return i;
    ops[nops] = Operand{src, ld, R, K, tr, bytes};
    bytes += gemm_hl_image_bytes(R, K);
    return nops++;
  };
  for (int i = 0; i < n; ++i) {
    a_of[i] = operand(d[i].A, d[i].lda, d[i].M, d[i].K, d[i].tA != 0);
    b_of[i] = operand(d[i].B, d[i].ldb, d[i].N, d[i].K, d[i].tB == 0);
  }
  if (bytes > w->g_cap) {
    if (w->g) {
      if ((e = cudaFree(w->g)) != cudaSuccess) return e;   // (waits for the device)
      w->g = nullptr;
      w->g_cap = 0;
    }
    if ((e = cudaMalloc((void**)&w->g, bytes)) != cudaSuccess) return e;
    w->g_cap = bytes;
  }
  static const int tiled_min = [] {
    const char* v = getenv("LSTMP_B200_SPLIT_TILED_MIN");
    return (v && *v) ? atoi(v) : 2 * 148;
  }();
  hl::SplitJobs sj;
  memset(&sj, 0, sizeof(sj));
  sj.n = nops;
  int nblk = 0;
  for (int i = 0; i < nops; ++i) {
    int blocks = 0;
    bool tiled = false;
    sj.j[i] = make_job(w->g + ops[i].off, ops[i].src, ops[i].ld, ops[i].R, ops[i].K, ops[i].transposed, &blocks, &tiled);
    if (ops[i].transposed && sj.j[i].nrt * sj.j[i].nkt >= tiled_min) {
      sj.j[i].transposed = 2;
      blocks = std::min(sj.j[i].nrt * sj.j[i].nkt, 148 * 8);
    }
    sj.blk0[i] = nblk;
    nblk += blocks;
  }
  for (int i = nops; i <= hl::MAX_SPLIT_JOBS; ++i) sj.blk0[i] = nblk;
  hl::split_multi_kernel<<<nblk, 256, 0, stream>>>(sj);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ++*nlaunch;

  // ---- split-K plan: cached per shape (a few thousand simulated schedules the first time)
  int key[1 + 3 * hl::MAX_GROUP] = {0};
  key[0] = n;
  for (int i = 0; i < n; ++i) {
    key[1 + 3 * i] = d[i].M;
    key[2 + 3 * i] = d[i].N;
    key[3 + 3 * i] = d[i].K;
  }
  HlWorkspace::GroupPlan* plan = nullptr;
  for (auto& pl : w->plans)
    if (memcmp(pl.key, key, sizeof(key)) == 0) plan = &pl;
  if (!plan) {
    HlWorkspace::GroupPlan np;
    memcpy(np.key, key, sizeof(key));
    hlplan::Shape shp[hl::MAX_GROUP];
    for (int i = 0; i < n; ++i) shp[i] = hlplan::Shape{d[i].M, d[i].N, d[i].K};
    hlplan::plan_group(shp, n, nsm, ws != nullptr, ws_floats, np.splits);
    w->plans.push_back(np);
    plan = &w->plans.back();
  }

  // ---- the products, longest items first
  int order[hl::MAX_GROUP];
  hl::HlProblem P[hl::MAX_GROUP];
  size_t ws_off = 0;
  for (int i = 0; i < n; ++i) {
    order[i] = i;
    P[i] = make_problem(d[i].C, d[i].ldc, d[i].M, d[i].N, d[i].K, d[i].alpha, w->g + ops[a_of[i]].off,
                        w->g + ops[b_of[i]].off, d[i].beta, d[i].bias, plan->splits[i], ws ? ws + ws_off : nullptr);
    if (P[i].splits > 1) ws_off += (size_t)P[i].splits * d[i].M * d[i].N;
  }
  std::stable_sort(order, order + n, [&](int x, int y) { return P[x].kt_per_split > P[y].kt_per_split; });
  hl::HlGroup g;
  memset(&g, 0, sizeof(g));
  g.n = n;
  hl::ReduceJobs rj;
  memset(&rj, 0, sizeof(rj));
  int rblk = 0;
  for (int k = 0; k < n; ++k) {
    const int i = order[k];
    g.p[k] = P[i];
    g.p[k].item0 = g.nitems;
    g.nitems += P[i].ntm * P[i].ntn * P[i].splits;
    if (P[i].splits > 1) {
      hl::ReduceJob& r = rj.j[rj.n];
      r.C = d[i].C; r.ws = P[i].split_ws; r.bias = d[i].bias; r.ldc = d[i].ldc;
      r.M = d[i].M; r.N = d[i].N; r.splits = P[i].splits; r.alpha = d[i].alpha; r.beta = d[i].beta;
      int rb = (int)(((long long)d[i].M * (d[i].N >> 2) + 255) / 256);
      rb = std::max(1, std::min(rb, 148 * 4));
      rj.blk0[rj.n++] = rblk;
      rblk += rb;
    }
  }
  for (int i = rj.n; i <= hl::MAX_GROUP; ++i) rj.blk0[i] = rblk;
  dim3 grid(g.nitems < nsm ? g.nitems : nsm), block(hl::THREADS);
  hl::gemm_hl_kernel<<<grid, block, hl::SMEM, stream>>>(g);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  ++*nlaunch;
  if (rj.n > 0) {
    hl::splitk_reduce_multi_kernel<<<rblk, 256, 0, stream>>>(rj);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*nlaunch;
  }
  *handled = true;
  return cudaSuccess;
}

void gemm_hl_free(HlWorkspace* w) {
  if (!w) return;
  if (w->a) cudaFree(w->a);
  if (w->b) cudaFree(w->b);
  if (w->g) cudaFree(w->g);
  w->a = w->b = w->g = nullptr;
  w->a_cap = w->b_cap = w->a_bytes = w->g_cap = 0;
  w->plans.clear();
}

}  // namespace lstmp
