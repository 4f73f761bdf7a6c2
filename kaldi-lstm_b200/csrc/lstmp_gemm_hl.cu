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
//   2. gemm_hl_kernel: one thread issues four 16 KB cp.async.bulk copies per 64-deep K block straight into the UMMA
//      ring (mbarrier expect_tx), one warp issues 16 MMAs per block, four warps run the epilogue.  No loader warps, no
//      register pass, no generic-proxy stores into operand tiles, no bounds logic on the operand side.
//
// (lstmp_gemm_tc.cu moves 32 KB in, 32 KB back out and 64 KB of hi/lo tiles through shared memory per 32-deep K block
// with eight loader warps and reaches ~1.9 k cycles per block against 0.78 k of MMA time)
#include <cuda_bf16.h>
#include <stdlib.h>

#include "lstmp_common.cuh"
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
__global__ void __launch_bounds__(256) split_hl_transposed_tiled_kernel(uint8_t* __restrict__ img,
                                                                        const float* __restrict__ src, long long ld,
                                                                        int R, int K, int nrt, int nkt) {
  __shared__ float tile[BK][129];
  const int rt = blockIdx.x % nrt, kt = blockIdx.x / nrt;
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

// PERSISTENT: grid = min(work items, #SMs) CTAs; work item w = (m tile, n tile, K split), m fastest (the CTAs that run
// at the same time share the B tile in L2), item w of CTA c = c + i * gridDim.x.  The accumulator is double-buffered in
// TMEM (2 x 128 columns): the four epilogue warps drain tile i (tcgen05.ld -> alpha/beta/bias -> 64 contiguous bytes
// per lane and load) while the producer / MMA warps are already in the main loop of tile i+1; the shared-memory ring
// runs continuously across tiles.  a_img / b_img: tile images of op(A) [M x K] and op(B)^T [N x K].
__global__ void __launch_bounds__(THREADS, 1)
gemm_hl_kernel(float* __restrict__ Cout, long long ldc_out, int M, int N, int nkt, float alpha_in,
               const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, float beta_in,
               const float* __restrict__ bias_in, int kt_per_split, int splits, float* __restrict__ split_ws,
               float* __restrict__ Cout2, long long ldc_out2, int N2) {
  // Cout2 != nullptr: TWO products that share op(A) in one launch -- C = A * B1 and C2 = A * B2 -- with the n tiles of
  // B2's image following those of B1 in b_img (the two weight-gradient GEMMs that both contract DGIFO^T).
  extern __shared__ __align__(128) uint8_t smem_raw_hl[];
  uint8_t* tiles = smem_raw_hl + ((1024u - (smem_u32(smem_raw_hl) & 1023u)) & 1023u);
  float* patches = reinterpret_cast<float*>(tiles + NSTAGE * STAGE);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + NSTAGE * STAGE + NEPI * PATCH);
  uint64_t* empty = full + NSTAGE;
  uint64_t* acc_full = empty + NSTAGE;   // [2] accumulator buffer complete (MMA warp -> epilogue)
  uint64_t* acc_empty = acc_full + 2;    // [2] accumulator buffer drained (4 epilogue warps -> MMA warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int ntn1 = (N + BN - 1) / BN;
  const int ntm = (M + BM - 1) / BM, ntn = ntn1 + (Cout2 ? (N2 + BN - 1) / BN : 0);
  const int nitems = ntm * ntn * splits;
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
        const int mt = w % ntm, nt = (w / ntm) % ntn, z = w / (ntm * ntn);
        const int kt0 = z * kt_per_split, nk = min(kt_per_split, nkt - kt0);
        const uint8_t* ap = a_img + ((size_t)mt * nkt + kt0) * 2 * TILE;   // A_hi | A_lo of (mt, kt) are adjacent
        const uint8_t* bp = b_img + ((size_t)nt * nkt + kt0) * 2 * TILE;
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
      const int z = w / (ntm * ntn);
      const int kt0 = z * kt_per_split, nk = min(kt_per_split, nkt - kt0);
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
    const int N_first = N;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < nitems; w += gridDim.x, ++it) {
      const int mt = w % ntm, nt = (w / ntm) % ntn, z = w / (ntm * ntn);
      const bool second = nt >= ntn1;            // (dual launch) this n tile belongs to the second product
      const int m0 = mt * BM, n0 = (second ? nt - ntn1 : nt) * BN;
      const int N = second ? N2 : N_first;
      float* Cm = second ? Cout2 : Cout;
      long long ldc = second ? ldc_out2 : ldc_out;
      float alpha = alpha_in, beta = beta_in;
      const float* bias = second ? nullptr : bias_in;
      if (split_ws) {  // split-K: raw partial sums to split_ws[z][M x N]; splitk_reduce applies alpha / beta / bias
        Cm = split_ws + (size_t)z * M * N;
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

cudaError_t gemm_hl_run(float* C, long long ldc, int M, int N, int K, float alpha, const uint8_t* a_img,
                        const uint8_t* b_img, float beta, const float* bias, cudaStream_t stream, float* ws,
                        size_t ws_floats, int* nlaunch, float* C2 = nullptr, long long ldc2 = 0, int N2 = 0) {
  *nlaunch = 1;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!attr_set[dev & 63]) {
    e = cudaFuncSetAttribute((const void*)hl::gemm_hl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hl::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  const int nkt = (K + hl::BK - 1) / hl::BK;
  const int ntm = (M + hl::BM - 1) / hl::BM,
            ntn = (N + hl::BN - 1) / hl::BN + (C2 ? (N2 + hl::BN - 1) / hl::BN : 0);
  const int tiles = ntm * ntn;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  int splits = 1;
  // split-K when the output tiles alone cannot fill the SMs (in_diff: 40 tiles, G(w_r_m): 28 tiles); never for a dual launch
  if (!C2 && ws && tiles < 100 && nkt >= 4 && (N & 3) == 0) {
    splits = nsm / tiles;
    if (splits > 8) splits = 8;
    if (splits > nkt / 2) splits = nkt / 2;
    while (splits > 1 && (size_t)splits * M * N > ws_floats) --splits;
  }
  int kts = nkt;
  if (splits > 1) {
    kts = (nkt + splits - 1) / splits;
    splits = (nkt + kts - 1) / kts;
  }
  const int items = tiles * splits;
  dim3 grid(items < nsm ? items : nsm), block(hl::THREADS);
  hl::gemm_hl_kernel<<<grid, block, hl::SMEM, stream>>>(C, ldc, M, N, nkt, alpha, a_img, b_img, beta, bias, kts, splits,
                                                        splits > 1 ? ws : nullptr, C2, ldc2, N2);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (splits > 1) {
    *nlaunch = 2;
    return launch_splitk_reduce(C, ldc, M, N, alpha, beta, bias, ws, splits, stream);
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

// C1[M x N1] = op(A) * B1 and C2[M x N2] = op(A) * B2 in ONE launch (alpha = 1, beta = 0, no bias): op(A) is split
// once, B1 and B2 (both stored [K x N], n contiguous) are split into consecutive n tiles of one image.  The two
// weight-gradient GEMMs of a layer that contract DGIFO^T: G(w_gifo_x) = DGIFO^T * in, G(w_gifo_r) = DGIFO^T * R.
cudaError_t launch_gemm_hl_dual(HlWorkspace* w, int M, int K, const float* A, long long lda, int tA, float* C1,
                                long long ldc1, int N1, const float* B1, long long ldb1, float* C2, long long ldc2, int N2,
                                const float* B2, long long ldb2, cudaStream_t stream, bool* handled, int* nlaunch) {
  *handled = false;
  *nlaunch = 0;
  if (!w || M <= 0 || N1 <= 0 || N2 <= 0 || K <= 0) return cudaSuccess;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(C1) || !al16(C2) || !al16(A) || !al16(B1) || !al16(B2)) return cudaSuccess;
  const size_t na = gemm_hl_image_bytes(M, K), nb1 = gemm_hl_image_bytes(N1, K), nb2 = gemm_hl_image_bytes(N2, K);
  auto grow = [&](uint8_t** p, size_t* cap, size_t need) -> cudaError_t {
    if (need <= *cap) return cudaSuccess;
    if (*p) {
      cudaError_t e = cudaFree(*p);
      if (e != cudaSuccess) return e;
      *p = nullptr;
      *cap = 0;
    }
    cudaError_t e = cudaMalloc((void**)p, need);
    if (e == cudaSuccess) *cap = need;
    return e;
  };
  cudaError_t e;
  int b1 = 0, b2 = 0;
  bool t1 = false, t2 = false;
  make_job(nullptr, B1, ldb1, N1, K, true, &b1, &t1);
  make_job(nullptr, B2, ldb2, N2, K, true, &b2, &t2);
  if (t1 || t2) return cudaSuccess;  // operands large enough for the tiled transposed split: two ordinary GEMMs
  if ((e = grow(&w->a, &w->a_cap, na)) != cudaSuccess) return e;
  if ((e = grow(&w->b, &w->b_cap, nb1 + nb2)) != cudaSuccess) return e;
  if ((e = gemm_hl_split(w->a, A, lda, M, K, tA != 0, stream)) != cudaSuccess) return e;
  w->a_bytes = na;
  const hl::SplitJob j1 = make_job(w->b, B1, ldb1, N1, K, true, &b1, &t1);
  const hl::SplitJob j2 = make_job(w->b + nb1, B2, ldb2, N2, K, true, &b2, &t2);
  hl::split_pair_kernel<<<b1 + b2, 256, 0, stream>>>(j1, j2, b1);   // (direct kernel for both: these operands are small)
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  int nl = 0;
  e = gemm_hl_run(C1, ldc1, M, N1, K, 1.f, w->a, w->b, 0.f, nullptr, stream, nullptr, 0, &nl, C2, ldc2, N2);
  *nlaunch = 2 + nl;
  *handled = true;
  return e;
}

void gemm_hl_free(HlWorkspace* w) {
  if (!w) return;
  if (w->a) cudaFree(w->a);
  if (w->b) cudaFree(w->b);
  w->a = w->b = nullptr;
  w->a_cap = w->b_cap = w->a_bytes = 0;
}

}  // namespace lstmp
