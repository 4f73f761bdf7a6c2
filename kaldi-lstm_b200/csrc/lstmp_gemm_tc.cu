// tcgen05 (5th-gen tensor core) GEMM with FP32-faithful 3xTF32 split arithmetic, sm_100a only.
//
//   C[M x N] = alpha * op(A) * op(B) + beta * C (+ bias[n])       fp32 in, fp32 out
//
// Used for the contractions OUTSIDE the time loop (CuMatrixBase::AddMatMat call sites
// LPS.h:246, :457, :468, :471, :486 of the reference), which are the dense stream-batched GEMMs.
//
// Precision: the path must match the reference's fp32 SGEMM to 1e-4, and single-pass TF32
// (10-bit mantissa) does not.  Every fp32 operand x is split on the fly into
//     hi = x & 0xffffe000   (exactly representable in TF32)
//     lo = x - hi           (exact in fp32; its own TF32 truncation error is ~2^-21 |x|)
// and the product is accumulated in FP32 in TMEM as  lo_a*hi_b + hi_a*lo_b + hi_a*hi_b.
//
// Structure of one CTA (288 threads), one 128 x BN output tile:
//   warps 0-7  : loader/transform -- cp.async.cg (LDGSTS) of the fp32 operands (either storage order) into per-thread
//                landing slots, 3 K blocks in flight; each thread reads its own 16-byte units back, transposes
//                m/n-contiguous sources in registers, splits into hi/lo and stores them (STS.128) into SWIZZLE_128B
//                K-major tiles; one mbarrier arrival per warp.  After the main loop the same warps run the epilogue:
//                tcgen05.ld -> alpha/beta/bias -> STG.
//   warp 8     : TMEM allocation; after its acquire-wait it executes the fence.proxy.async for the loaders' stores and
//                issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) x 12 per 32-deep K block (whole warp on warp-uniform
//                operands, elect.sync predication) and tcgen05.commit onto the stage's "empty" mbarrier / the
//                accumulator-ready mbarrier.
// 2-stage UMMA ring + 3 raw landing slots, mbarrier full/empty handshakes, accumulator in TMEM.
// Since round 2 this kernel is the FALLBACK (LSTMP_B200_GEMM=1); the default is lstmp_gemm_hl.cu (pre-split bf16 hi/lo
// tile images, bulk copies, no loader warps).
#include <cstdlib>
#include "lstmp_common.cuh"
#include "lstmp_kernels.h"
#include "lstmp_tc.cuh"

namespace lstmp {

namespace tc {
constexpr int BM = 128;
constexpr int BK = 32;        // fp32 elements per stage along K (= 4 MMAs of K=8)
constexpr int NRAW = 3;        // K blocks of raw fp32 operands in flight per CTA (cp.async landing slots)
constexpr int LOADERS = 256;  // warps 0-7
constexpr int THREADS = LOADERS + 32;

// operand tiles are SWIZZLE_128B K-major (lstmp_tc.cuh): rows x 128 bytes, dense
__host__ __device__ constexpr uint32_t tile_bytes(int rows, bool) { return (uint32_t)rows * 128u; }

// Global operand load that the compiler may not sink towards its use (register-prefetch variant).
__device__ __forceinline__ float4 ldg_early(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// One operand's loader state.  The shared-memory tile is ALWAYS K-major (the layout validated on
// hardware); a source stored with the M/N index contiguous ("MN-major", e.g. DGIFO^T for the weight
// gradients) is transposed on the fly in registers: each thread owns a 4(k) x 4(mn) block, loads it with
// four coalesced LDG.128 along mn and stores four 16-byte K-chunks, one per mn row.
template <int ROWS, bool MN, int NT = LOADERS>  // NT = threads that share one K block's tile (tid = index among them)
struct Loader {
  static constexpr int UNITS = ROWS * (BK / 4) / NT;              // float4 units per thread (K-major source)
  static constexpr int NBLK = (ROWS / 4) * (BK / 4);                    // 4x4 blocks per tile (MN-major source)
  static constexpr int BPT = (NBLK + NT - 1) / NT;            // blocks per thread
  static constexpr int NV = MN ? BPT * 4 : UNITS;                       // 16-byte units this thread owns per K block
  static constexpr uint32_t RAW_BYTES = (uint32_t)NV * NT * 16;    // raw landing slot of one K block
  float4 v[NV];

  // Stage 1: cp.async (LDGSTS, tracked by cp.async groups, NRAW K blocks in flight) of this thread's units of the K
  // block at k0 into ITS OWN 16-byte slots of `raw`.  (Measured against prefetching into rotating register sets, the
  // load() variant below: 43 vs 52 us on the weight-gradient GEMMs.)
  __device__ __forceinline__ void issue(uint8_t* raw, const float* __restrict__ src, long long ld, int row0,
                                        int nrows, int k0, int K, int tid) const {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!MN) {
#pragma unroll
      for (int i = 0; i < UNITS; ++i) {
        const int u = tid + i * NT;
        const int kc = u % (BK / 4), r = u / (BK / 4);
        const int gr = row0 + r, gk = k0 + 4 * kc;
        uint8_t* dst = raw + (size_t)u * 16;
        if (gr < nrows && gk < K) cp_async16(dst, src + (size_t)gr * ld + gk);
        else *reinterpret_cast<float4*>(dst) = z;
      }
    } else {
      // blocks of 4 k x 4 mn: block b -> mn-block mb = b % (ROWS/4), k-chunk kc = b / (ROWS/4)
#pragma unroll
      for (int blk = 0; blk < BPT; ++blk) {
        const int b = tid + blk * NT;
        const int mb = b % (ROWS / 4), kc = b / (ROWS / 4);
        const int gr = row0 + 4 * mb;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int gk = k0 + 4 * kc + kk;
          uint8_t* dst = raw + ((size_t)(blk * 4 + kk) * NT + tid) * 16;
          if (b < NBLK && gr < nrows && gk < K) cp_async16(dst, src + (size_t)gk * ld + gr);
          else *reinterpret_cast<float4*>(dst) = z;  // row kk of the block: 4 consecutive mn at k = 4kc+kk
        }
      }
    }
  }
  // Register-prefetch variant (STAGED = false): the same units straight into v[] with LDG.128; the caller rotates
  // several Loader objects so that the loads of later K blocks are in flight.  Half the shared-memory traffic of the
  // staged variant (no landing slot to write and read back).
  __device__ __forceinline__ void load(const float* __restrict__ src, long long ld, int row0, int nrows, int k0,
                                       int K, int tid) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!MN) {
#pragma unroll
      for (int i = 0; i < UNITS; ++i) {
        const int u = tid + i * NT;
        const int kc = u % (BK / 4), r = u / (BK / 4);
        const int gr = row0 + r, gk = k0 + 4 * kc;
        v[i] = (gr < nrows && gk < K) ? ldg_early(src + (size_t)gr * ld + gk) : z;
      }
    } else {
#pragma unroll
      for (int blk = 0; blk < BPT; ++blk) {
        const int b = tid + blk * NT;
        const int mb = b % (ROWS / 4), kc = b / (ROWS / 4);
        const int gr = row0 + 4 * mb;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int gk = k0 + 4 * kc + kk;
          v[blk * 4 + kk] = (b < NBLK && gr < nrows && gk < K) ? ldg_early(src + (size_t)gk * ld + gr) : z;
        }
      }
    }
  }
  // Stage 2a: read this thread's own units back (after cp.async.wait_group)
  __device__ __forceinline__ void fetch(const uint8_t* raw, int tid) {
    if (!MN) {
#pragma unroll
      for (int i = 0; i < UNITS; ++i) v[i] = *reinterpret_cast<const float4*>(raw + (size_t)(tid + i * NT) * 16);
    } else {
#pragma unroll
      for (int i = 0; i < BPT * 4; ++i) v[i] = *reinterpret_cast<const float4*>(raw + ((size_t)i * NT + tid) * 16);
    }
  }

  static __device__ __forceinline__ void split_store(uint8_t* hi, uint8_t* lo, uint32_t off, float4 x) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
    l.x = x.x - h.x;
    l.y = x.y - h.y;
    l.z = x.z - h.z;
    l.w = x.w - h.w;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }

  __device__ __forceinline__ void store(uint8_t* hi, uint8_t* lo, int tid) const {
    if (!MN) {
#pragma unroll
      for (int i = 0; i < UNITS; ++i) {
        const int u = tid + i * NT;
        const int kc = u % (BK / 4), r = u / (BK / 4);
        split_store(hi, lo, sw128_off(r, kc), v[i]);
      }
    } else {
#pragma unroll
      for (int blk = 0; blk < BPT; ++blk) {
        const int b = tid + blk * NT;
        if (b >= NBLK) break;
        const int mb = b % (ROWS / 4), kc = b / (ROWS / 4);
        const float4 r0 = v[blk * 4 + 0], r1 = v[blk * 4 + 1], r2 = v[blk * 4 + 2], r3 = v[blk * 4 + 3];
        // transposed columns: c_i = (k0..k3) at mn = 4mb+i
        const float4 c0 = make_float4(r0.x, r1.x, r2.x, r3.x);
        const float4 c1 = make_float4(r0.y, r1.y, r2.y, r3.y);
        const float4 c2 = make_float4(r0.z, r1.z, r2.z, r3.z);
        const float4 c3 = make_float4(r0.w, r1.w, r2.w, r3.w);
        // rotate the store order per lane so the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups
        const int rot = (mb >> 1) & 3;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int i = (t + rot) & 3;
          const float4 c = (i == 0) ? c0 : (i == 1) ? c1 : (i == 2) ? c2 : c3;
          const int r = 4 * mb + i;
          split_store(hi, lo, sw128_off(r, kc), c);
        }
      }
    }
  }
};

// MODE: 1 = cp.async landing slots (the only mode instantiated: 43 vs 52 us on the weight-gradient GEMMs against MODE 0,
//       LDG.128 register prefetch; a "ping-pong" two-group loader was measured slower still in round 2 and removed)
template <int BN, bool A_MN, bool B_MN, int MODE>
struct Smem {
  static constexpr bool STAGED = (MODE == 1);
  static constexpr int NSTAGE = STAGED ? 2 : 3;                  // UMMA operand stages (hi/lo tiles of A and B)
  static constexpr uint32_t A_BYTES = tile_bytes(BM, false);
  static constexpr uint32_t B_BYTES = tile_bytes(BN, false);
  static constexpr uint32_t STAGE = 2 * A_BYTES + 2 * B_BYTES;  // A_hi | A_lo | B_hi | B_lo
  static constexpr uint32_t RAW_A = Loader<BM, A_MN>::RAW_BYTES, RAW_B = Loader<BN, B_MN>::RAW_BYTES;
  static constexpr uint32_t RAW = RAW_A + RAW_B;                 // one K block of raw fp32 operands
  static constexpr uint32_t TOTAL = NSTAGE * STAGE + (STAGED ? NRAW * RAW : 0) + 2048;  // + barriers, alignment slack
};

template <int BN, bool A_MN, bool B_MN, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(float* __restrict__ Cm, long long ldc, int M, int N, int K, float alpha,
               const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb, float beta,
               const float* __restrict__ bias, int k_per_split, float* __restrict__ split_ws) {
  using SM = Smem<BN, A_MN, B_MN, MODE>;
  constexpr bool STAGED = SM::STAGED;
  constexpr int NSTAGE = SM::NSTAGE;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // 1024-byte align the tile area (SWIZZLE_128B atoms)
  // Pointer arithmetic on the __shared__ array (not an integer round trip) keeps the address space visible to the
  // compiler: the loaders' accesses are STS.128 / LDS.128 instead of generic ST.E / LD.E (round 2: 47 vs 50 us in_diff)
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* raw = tiles + NSTAGE * SM::STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(raw + (STAGED ? NRAW * SM::RAW : 0));
  uint64_t* empty = full + NSTAGE;
  uint64_t* accum_ready = empty + NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_ready + 1);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform (keeps the issuer's operands in URs)
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (split_ws) {
    // split-K: slice z of the K range; raw partial sums go to split_ws[z][M x N] and a second kernel reduces them
    // in a fixed order (deterministic), applying alpha / beta / bias there.
    const int kb0 = blockIdx.z * k_per_split;
    A += A_MN ? (size_t)kb0 * lda : (size_t)kb0;
    B += B_MN ? (size_t)kb0 * ldb : (size_t)kb0;
    K = (K - kb0 < k_per_split) ? K - kb0 : k_per_split;
    Cm = split_ws + (size_t)blockIdx.z * M * N;
    ldc = N;
    alpha = 1.f;
    beta = 0.f;
    bias = nullptr;
  }
  const int nkb = (K + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], LOADERS / 32);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_ready, 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    // TMEM: BN fp32 accumulator columns x 128 lanes (power of two >= 32 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "n"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < 8) {
    // =============================== loader / transform ==================================
    if constexpr (STAGED) {
      Loader<BM, A_MN> la;
      Loader<BN, B_MN> lb;
  #pragma unroll
      for (int i = 0; i < NRAW; ++i) {
        if (i < nkb) {
          la.issue(raw + i * SM::RAW, A, lda, m0, M, i * BK, K, tid);
          lb.issue(raw + i * SM::RAW + SM::RAW_A, B, ldb, n0, N, i * BK, K, tid);
        }
        cp_async_commit();
      }
      int rs = 0;  // raw slot of K block kb
      for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<NRAW - 1>();  // this thread's copies of K block kb have landed
        uint8_t* rw = raw + rs * SM::RAW;
        la.fetch(rw, tid);
        lb.fetch(rw + SM::RAW_A, tid);
        if (kb + NRAW < nkb) {
          la.issue(rw, A, lda, m0, M, (kb + NRAW) * BK, K, tid);
          lb.issue(rw + SM::RAW_A, B, ldb, n0, N, (kb + NRAW) * BK, K, tid);
        }
        cp_async_commit();  // one group per K block (possibly empty) keeps the wait_group arithmetic uniform
        if (++rs == NRAW) rs = 0;
        const int s = kb % NSTAGE;
        if (kb >= NSTAGE) mbar_wait(&empty[s], (uint32_t)(((kb / NSTAGE) - 1) & 1));
        uint8_t* st = tiles + (size_t)s * SM::STAGE;
        la.store(st, st + SM::A_BYTES, tid);
        lb.store(st + 2 * SM::A_BYTES, st + 2 * SM::A_BYTES + SM::B_BYTES, tid);
        // The proxy fence lives on the consumer side (MMA warp, after its acquire-wait): here it would compile to
        // MEMBAR.ALL.CTA and wait for the copies of the next K blocks.  One arrival per warp.
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&full[s]);
      }
    } else {
      // Four register sets per operand: the global loads of K blocks kb+1 .. kb+3 are in flight while block kb is
      // split and stored.
      constexpr int PF = 4;
      Loader<BM, A_MN> la0, la1, la2, la3;
      Loader<BN, B_MN> lb0, lb1, lb2, lb3;
      la0.load(A, lda, m0, M, 0, K, tid);
      lb0.load(B, ldb, n0, N, 0, K, tid);
      if (nkb > 1) { la1.load(A, lda, m0, M, BK, K, tid); lb1.load(B, ldb, n0, N, BK, K, tid); }
      if (nkb > 2) { la2.load(A, lda, m0, M, 2 * BK, K, tid); lb2.load(B, ldb, n0, N, 2 * BK, K, tid); }
      if (nkb > 3) { la3.load(A, lda, m0, M, 3 * BK, K, tid); lb3.load(B, ldb, n0, N, 3 * BK, K, tid); }
      auto step = [&](Loader<BM, A_MN>& la, Loader<BN, B_MN>& lb, int kb) {
        const int s = kb % NSTAGE;
        if (kb >= NSTAGE) mbar_wait(&empty[s], (uint32_t)(((kb / NSTAGE) - 1) & 1));
        uint8_t* st = tiles + (size_t)s * SM::STAGE;
        la.store(st, st + SM::A_BYTES, tid);
        lb.store(st + 2 * SM::A_BYTES, st + 2 * SM::A_BYTES + SM::B_BYTES, tid);
        if (kb + PF < nkb) {
          la.load(A, lda, m0, M, (kb + PF) * BK, K, tid);
          lb.load(B, ldb, n0, N, (kb + PF) * BK, K, tid);
        }
        __syncwarp();  // proxy fence on the consumer side; one arrival per warp
        if ((tid & 31) == 0) mbar_arrive(&full[s]);
      };
      for (int kb = 0; kb < nkb; kb += PF) {
        step(la0, lb0, kb);
        if (kb + 1 < nkb) step(la1, lb1, kb + 1);
        if (kb + 2 < nkb) step(la2, lb2, kb + 2);
        if (kb + 3 < nkb) step(la3, lb3, kb + 3);
      }
    }
    // =============================== epilogue ============================================
    mbar_wait(accum_ready, 0);
    tc_fence_after();
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = warp >> 2;           // column half
    const int lane = tid & 31;
    const int gm = m0 + quad * 32 + lane;
    constexpr int COLS_PER_WARP = BN / 2;
#pragma unroll 1
    for (int c = 0; c < COLS_PER_WARP; c += 16) {
      const int col = half * COLS_PER_WARP + c;
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col, v);
      if (gm < M) {
        float* crow = Cm + (size_t)gm * ldc;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int gn = n0 + col + j;
          if (gn + 3 < N && ((ldc & 3) == 0)) {
            float4 o = make_float4(alpha * v[j], alpha * v[j + 1], alpha * v[j + 2], alpha * v[j + 3]);
            if (beta != 0.f) {
              float4 cc = *reinterpret_cast<const float4*>(crow + gn);
              o.x += beta * cc.x; o.y += beta * cc.y; o.z += beta * cc.z; o.w += beta * cc.w;
            }
            if (bias) {
              float4 bb = __ldg(reinterpret_cast<const float4*>(bias + gn));
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            *reinterpret_cast<float4*>(crow + gn) = o;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (gn + q < N) {
                float o = alpha * v[j + q];
                if (beta != 0.f) o += beta * crow[gn + q];
                if (bias) o += bias[gn + q];
                crow[gn + q] = o;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // =============================== MMA issuer (warp 8) ==================================
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
    // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);  // both operands K-major in shared memory
    const uint32_t tiles_s = smem_u32(tiles);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % NSTAGE;
      mbar_wait(&full[s], (uint32_t)((kb / NSTAGE) & 1));
      fence_async_smem();  // loaders' generic-proxy stores (ordered by the mbarrier) -> tensor core's async proxy
      tc_fence_after();
      {
        // the whole warp runs the burst on warp-uniform operands; elect.sync predicates the instructions (lstmp_tc.cuh)
        const uint32_t a_hi = tiles_s + s * SM::STAGE, a_lo = a_hi + SM::A_BYTES;
        const uint32_t b_hi = a_hi + 2 * SM::A_BYTES, b_lo = b_hi + SM::B_BYTES;
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          // MMA j covers the 16-byte K chunks 2j and 2j+1 of the 128-byte swizzled rows (K = 8 tf32 per instruction)
          const uint64_t dah = make_desc_sw128(a_hi + 32 * j), dal = make_desc_sw128(a_lo + 32 * j);
          const uint64_t dbh = make_desc_sw128(b_hi + 32 * j), dbl = make_desc_sw128(b_lo + 32 * j);
          if (elect_one()) mma_tf32(tmem_base, dal, dbh, idesc, (kb | j) ? 1u : 0u);  // small terms first
          if (elect_one()) mma_tf32(tmem_base, dah, dbl, idesc, 1u);
          if (elect_one()) mma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        if (elect_one()) {
          umma_commit(&empty[s]);                       // frees the stage once these MMAs have read it
          if (kb == nkb - 1) umma_commit(accum_ready);  // accumulator complete
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(float* __restrict__ C, long long ldc, int M, int N,
                                                            float alpha, float beta, const float* __restrict__ bias,
                                                            const float* __restrict__ ws, int splits) {
  const int n4 = N >> 2;
  const long long total = (long long)M * n4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / n4), n = (int)(idx - (long long)m * n4) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
      const float4 v = *reinterpret_cast<const float4*>(ws + ((size_t)z * M + m) * N + n);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    float* c = C + (size_t)m * ldc + n;
    float o[4] = {alpha * a.x, alpha * a.y, alpha * a.z, alpha * a.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (beta != 0.f) o[q] += beta * c[q];
      if (bias) o[q] += bias[n + q];
      c[q] = o[q];
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int MODE>
static cudaError_t launch_one(float* C, long long ldc, int M, int N, int K, float alpha, const float* A,
                              long long lda, const float* B, long long ldb, float beta, const float* bias,
                              cudaStream_t stream, float* ws, size_t ws_floats, int* nlaunch) {
  using SM = Smem<BN, A_MN, B_MN, MODE>;
  *nlaunch = 1;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!attr_set[dev & 63]) {
    e = cudaFuncSetAttribute((const void*)gemm_tc_kernel<BN, A_MN, B_MN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SM::TOTAL);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM), block(THREADS);
  // split-K when the output tiles alone cannot fill the 148 SMs (e.g. in_diff: 40 tiles, G(w_r_m): 28 tiles)
  const int tiles = grid.x * grid.y, nkb = (K + BK - 1) / BK;
  int splits = 1;
  if (ws && tiles < 100 && nkb >= 8 && (N & 3) == 0) {
    splits = 148 / tiles;
    if (splits > 8) splits = 8;
    if (splits > nkb / 4) splits = nkb / 4;
    while (splits > 1 && (size_t)splits * M * N > ws_floats) --splits;
  }
  if (splits > 1) {
    const int kbs = (nkb + splits - 1) / splits;        // K blocks per split
    splits = (nkb + kbs - 1) / kbs;                      // drop empty tail splits
    grid.z = splits;
    gemm_tc_kernel<BN, A_MN, B_MN, MODE><<<grid, block, SM::TOTAL, stream>>>(C, ldc, M, N, K, alpha, A, lda, B, ldb, beta,
                                                                       bias, kbs * BK, ws);
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return e2;
    int rb = (int)(((long long)M * (N >> 2) + 255) / 256);
    if (rb > 148 * 4) rb = 148 * 4;
    splitk_reduce_kernel<<<rb, 256, 0, stream>>>(C, ldc, M, N, alpha, beta, bias, ws, splits);
    *nlaunch = 2;
    return cudaGetLastError();
  }
  gemm_tc_kernel<BN, A_MN, B_MN, MODE><<<grid, block, SM::TOTAL, stream>>>(C, ldc, M, N, K, alpha, A, lda, B, ldb, beta,
                                                                     bias, 0, nullptr);
  return cudaGetLastError();
}
}  // namespace tc

cudaError_t launch_splitk_reduce(float* C, long long ldc, int M, int N, float alpha, float beta, const float* bias,
                                 const float* ws, int splits, cudaStream_t stream) {
  int rb = (int)(((long long)M * (N >> 2) + 255) / 256);
  if (rb > 148 * 4) rb = 148 * 4;
  if (rb < 1) rb = 1;
  tc::splitk_reduce_kernel<<<rb, 256, 0, stream>>>(C, ldc, M, N, alpha, beta, bias, ws, splits);
  return cudaGetLastError();
}

// tA/tB as in launch_gemm: op(A) is M x K.  tA == 0: A stored [M x K] (k contiguous -> K-major);
// tA == 1: A stored [K x M] (m contiguous -> MN-major).  op(B) is K x N.  tB == 0: B stored [K x N]
// (n contiguous -> MN-major); tB == 1: B stored [N x K] (k contiguous -> K-major).
cudaError_t launch_gemm_tc(float* C, long long ldc, int M, int N, int K, float alpha, const float* A, long long lda,
                           int tA, const float* B, long long ldb, int tB, float beta, const float* bias,
                           cudaStream_t stream, bool* handled, float* ws, size_t ws_floats, int* nlaunch) {
  *handled = false;
  if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool a_mn = (tA != 0), b_mn = (tB == 0);
  // 128-bit loads: contiguous extent and leading dimension multiples of 4 floats, 16-byte bases
  if (!al16(A) || !al16(B) || (lda & 3) || (ldb & 3)) return cudaSuccess;
  if ((K & 3) && (!a_mn || !b_mn)) return cudaSuccess;
  if (a_mn && (M & 3)) return cudaSuccess;
  if (b_mn && (N & 3)) return cudaSuccess;
  if (bias && !al16(bias)) return cudaSuccess;
  if (!al16(C)) return cudaSuccess;
  *handled = true;
  const bool small_n = (N <= 64);
#define LSTMP_TC_CASE(BN_, AMN_, BMN_) \
  return tc::launch_one<BN_, AMN_, BMN_, 1>(C, ldc, M, N, K, alpha, A, lda, B, ldb, beta, bias, stream, ws, ws_floats, nlaunch)
  if (small_n) {
    if (!a_mn && !b_mn) LSTMP_TC_CASE(64, false, false);
    if (!a_mn && b_mn) LSTMP_TC_CASE(64, false, true);
    if (a_mn && !b_mn) LSTMP_TC_CASE(64, true, false);
    LSTMP_TC_CASE(64, true, true);
  } else {
    if (!a_mn && !b_mn) LSTMP_TC_CASE(128, false, false);
    if (!a_mn && b_mn) LSTMP_TC_CASE(128, false, true);
    if (a_mn && !b_mn) LSTMP_TC_CASE(128, true, false);
    LSTMP_TC_CASE(128, true, true);
  }
#undef LSTMP_TC_CASE
}

}  // namespace lstmp
