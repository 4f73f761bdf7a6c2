// tcgen05 time loops for LstmProjectedStreams on sm_100a, fed by bulk copies (the TMA unit's cp.async.bulk): forward
// (LPS.h:261-331) and the mirrored truncated-BPTT backward (LPS.h:369-454), one persistent launch per chunk and
// direction, <= 64 streams per stream group.  (LPS.h = google/nnet/bd-nnet-lstm-projected-streams.h of the reference.)
//
// Every per-timestep contraction is  D[128 x N] (TMEM, fp32) += A[128 x 16] * B[N x 16]^T  (tcgen05.mma.kind::f16,
// bf16 operands).  FP32 fidelity comes from a two-piece bf16 split  x = hi + lo  (hi = bf16(x), lo = bf16(x - hi))
// of BOTH operands, stacked instead of issued as extra instructions:
//   A rows 0..Sg-1  = hi halves of the all-gathered activations of the group's streams, rows Sg..2Sg-1 = lo halves;
//   B rows 0..n-1   = hi halves of the CTA's stationary weight slice,                   rows n..2n-1   = lo halves;
// one MMA yields all four cross products, result[s][j] = D[s][j] + D[s][n+j] + D[Sg+s][j] + D[Sg+s][n+j]
// (relative error ~4e-6 end to end, tools/split_precision_study.py; the path's tolerance is 1e-4).
//
// The PRODUCER of an activation splits it: the CTA that computes r(t) / m(t) / d_r(t) / DGIFO(t) writes the hi and lo
// bf16 halves straight into small global "tile image" arrays -- per stream group and 64-k chunk the ready-made
// SWIZZLE_128B K-major shared-memory image [hi rows | lo rows] x 128 bytes (store_hl), L2-resident -- and after the grid
// barrier every consumer pulls groups of consecutive tile images into a shared-memory ring with ONE cp.async.bulk each
// (SASS UBLKCP, mbarrier expect_tx).  The operands reach shared memory through the async proxy: no loader warps, no
// register pass, no generic stores into operand tiles, no proxy-fence relay -- up to three threads issue the copies,
// one warp issues the MMAs.  (Tensor-map copies, cp.async.bulk.tensor, were measured first: one 128-byte row per ~3.5
// cycles, 36-43 B/clk per SM -- DESIGN.md section 3.1.)  The grid barrier is split and run by ONE thread (GridSync):
// only the copies depend on other CTAs' data; the activation record goes to HBM after the arrive.
//
// Backward decomposition (the reason it differs from forward): d_r(t) = out_diff(t) + DGIFO(t+1) * W_gifo_r contracts
// over K = 4C.  CTAs form clusters of kp; cluster b owns ~R/np columns of d_r, CTA rank a of the cluster contracts the
// a-th K slice (it gathers 4C/kp columns of DGIFO(t+1) only) and the kp partial [Sg x R/np] blocks are summed through
// distributed shared memory (a cluster barrier, not a grid barrier).  d_m(t) = d_r(t) * W_r_m and the derivative
// chain run cell-sliced over all CTAs as in forward.  Two grid barriers per timestep in both directions.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "lstmp_common.cuh"
#include "lstmp_kernels.h"
#include "lstmp_tc.cuh"

namespace lstmp {

namespace tm {
constexpr int KC = 64;                       // bf16 k per ring slot row (128 bytes): 4 MMAs of K = 16
constexpr uint32_t TILE_MAX = 128 * 128;     // one tile: [<= 128 rows][128 B], SWIZZLE_128B; hi rows at 0, lo rows at row Sg
constexpr int NPROD = 3;                     // TMA producer warps (8, 10, 11): chunk c is issued by producer c mod 3
constexpr int MAX_SLOTS = 8;
constexpr uint32_t TMEM_COLS = 512;          // whole TMEM: the CTA is alone on its SM
constexpr uint32_t COL_A = 0, COL_B = 256;   // accumulator columns of the first / second product of a timestep

struct Pipe {
  uint32_t cc;   // ring chunks produced / consumed so far (same sequence in every thread)
  uint32_t acc;  // products finished so far (phase of the accumulator-ready barrier)
};

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
// kind::f16 with bf16 operands, fp32 accumulate, both operands K-major (cute::UMMA::InstrDescriptor):
// D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& h, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  l = __float2bfloat16_rn(x - __bfloat162float(h));
}
// 8 consecutive k of one weight row -> one 16-byte unit of the hi tile row and one of the lo tile row
__device__ __forceinline__ void split8_store(const float* v, uint8_t* hi_dst, uint8_t* lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(v[2 * i], h0, l0);
    split_bf16(v[2 * i + 1], h1, l1);
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_saddr, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];\n" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

// generic-proxy global stores <-> async-proxy (TMA) global reads: FENCE.VIEW.ASYNC.G only (the unqualified
// fence.proxy.async adds a MEMBAR.ALL.GPU); gpu-scope ordering itself comes from the grid barrier's release / acquire
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;\n" ::: "memory"); }

__device__ __forceinline__ void tma_bulk_g2s_u32(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar);

struct Ring {
  uint8_t* ring;
  uint64_t *full, *empty, *accum, *gridok;
  int nslot, nprod;
  int gt;               // tiles (64-k chunks) per ring slot = per bulk copy
  uint32_t slot_bytes;  // gt * tile bytes
};

// Grid barrier between the co-resident CTAs of one stream group, split into arrive / wait and run by ONE thread per
// CTA (lane 0 of the TMA producer warp): only the TMA reads depend on other CTAs' data, so nobody else ever waits for
// the grid -- the other warps are gated by the mbarrier chain full -> MMA -> accum.  Same monotonic counter as
// GroupBarrier (lstmp_common.cuh): barrier k is complete when counter - base >= k * nctas.  A CTA must wait for
// barrier k before it arrives at barrier k+1 (the same thread does both, in program order).
struct GridSync {
  unsigned* counter;
  unsigned target;
  unsigned nctas;
  bool off;
  // after a __syncthreads that ordered the CTA's hi/lo stores (each followed by fence.proxy.async.global)
  __device__ __forceinline__ void arrive() {
    target += nctas;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    stamp(200);
  }
  __device__ __forceinline__ void wait() {
    if (!off) {
      // acquire LOADS, not relaxed loads + fence.acq_rel.gpu: the fence also waits for this thread's own record
      // stores still in flight (measured ~1200 cycles per barrier); one polling thread per CTA makes the L1
      // invalidate that comes with every acquire load harmless
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(counter) : "memory");
      } while (static_cast<int>(v - target) < 0);
      stamp(201);
    }
    fence_async_global();  // other CTAs' generic-proxy stores before my async-proxy (TMA) reads
    stamp(202);
  }
};
constexpr int kProducerWarp = 8, kIssuerWarp = 9;
__device__ __forceinline__ bool is_sync_thread() { return threadIdx.x == kProducerWarp * 32; }

// red[row*ldred + n] = D[row][n] + D[row][nh + n] for the 128 stacked rows, n < nvalid, where
// D = A * B^T over K = 64*nch:  A = the tile images (kc0 + chunk) of the global exchange array `img` (each the
// ready-made SWIZZLE_128B shared-memory image of [hi rows | lo rows] x 64 k, written by the producers of the
// activation, see store_hl); B = the stationary tiles b_addr + chunk * chunk_b.  Chunks are walked in
// the rotated order chunk = (c + rot) mod nch (order-independent sum up to fp32 rounding, fixed per CTA:
// bit-reproducible; spreads the CTAs' requests for the same lines over time).  grid_wait: the operand was written by
// other CTAs before the grid barrier this CTA last arrived at -- the producer waits for it before its first copy.
// Every thread of the CTA calls this (CTA-uniform arguments); contains one __syncthreads at the end.
template <int GT>  // tiles per ring slot / bulk copy (compile time: the MMA burst of a slot is fully unrolled)
__device__ __forceinline__ void tma_product(Pipe& ps, const Ring& rg, GridSync& gs, bool grid_wait,
                                            const uint8_t* img, int Sg, int kc0, int nch, int rot,
                                            uint32_t b_addr, uint32_t chunk_b, uint32_t idesc, uint32_t tmem_d, int nh,
                                            int nvalid, float* red, int ldred) {
  // warp index through a shuffle: provably warp-uniform for ptxas (UMMA operands stay in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int nslot = rg.nslot;
  constexpr int gt = GT;
  const int ngr = (nch + gt - 1) / gt;  // pipeline steps: groups of up to gt consecutive tiles, one bulk copy each
  rot = rot % ngr;                       // the stagger rotates the walk over the groups
  const uint32_t tile_bytes = (uint32_t)(((2 * Sg + 7) & ~7) * 128);
  uint32_t slot = ps.cc % (uint32_t)nslot, use = ps.cc / (uint32_t)nslot;
  // With fewer ring slots than producers a producer could lap the parity of a slot's "empty" barrier (it would test
  // the phase two uses back): never more producers than slots.
  const int nprod = nslot < rg.nprod ? nslot : rg.nprod;
  const int prod = warp == kProducerWarp ? 0 : warp == kProducerWarp + 2 ? 1 : warp == kProducerWarp + 3 ? 2 : -1;
  if (prod >= 0 && prod < nprod) {
    // ------------------------------ TMA producers --------------------------------------------
    // One 1-D bulk copy (cp.async.bulk, SASS UBLKCP) per GROUP of gt tiles: the global array already holds the
    // swizzled tile images back to back.  A copy costs ~400 cycles of latency in the SM's copy unit whatever its size
    // (measured with 8 KB and 16 KB copies, and with tensor-map copies of 128-byte rows: 36-43 B/clk per SM,
    // profiles/r2_stamps_*_tensormap.txt), so fewer, larger copies are what helps; the groups are dealt round-robin to
    // up to three threads in three warps.
    if (lane == 0) {
      if (grid_wait) {
        if (prod == 0) {
          gs.wait();
          tc::mbar_arrive(rg.gridok);
        } else {
          mbar_wait(rg.gridok, ps.acc & 1);
        }
      }
      const uint32_t ring_s = smem_u32(rg.ring);
      for (int c = prod; c < ngr; c += nprod) {
        const uint32_t idx = ps.cc + (uint32_t)c, sl = idx % (uint32_t)nslot, us = idx / (uint32_t)nslot;
        if (us > 0) mbar_wait(&rg.empty[sl], (us - 1) & 1);
        const int ge = (c + rot < ngr) ? c + rot : c + rot - ngr;
        const int t0 = ge * gt, nt = min(gt, nch - t0);
        const uint32_t bytes = (uint32_t)nt * tile_bytes;
        mbar_arrive_expect_tx(&rg.full[sl], bytes);
        tma_bulk_g2s_u32(ring_s + sl * rg.slot_bytes, img + (size_t)(kc0 + t0) * tile_bytes, bytes, &rg.full[sl]);
      }
      stamp(203);
    }
    __syncwarp();
  } else if (warp == kIssuerWarp) {
    // ------------------------------ MMA issuer -----------------------------------------------
    const uint32_t ring_s = smem_u32(rg.ring);
    for (int c = 0; c < ngr; ++c) {
      mbar_wait(&rg.full[slot], use & 1);
      if (c == 0) stamp(210);
      tc::tc_fence_after();
      {
        // the whole warp runs the burst on warp-uniform operands; elect.sync predicates the instructions (lstmp_tc.cuh)
        const int ge = (c + rot < ngr) ? c + rot : c + rot - ngr;
        const int t0 = ge * gt, nt = min(gt, nch - t0);
#pragma unroll
        for (int t = 0; t < GT; ++t) {
          if (t < nt) {  // (warp-uniform; only the last group of a product can be short)
            const uint32_t a0 = ring_s + slot * rg.slot_bytes + (uint32_t)t * tile_bytes;
            const uint32_t b0 = b_addr + (uint32_t)(t0 + t) * chunk_b;
#pragma unroll
            for (int j = 0; j < KC / 16; ++j) {
              // MMA j covers the 16-byte units 2j, 2j+1 of the 128-byte rows (K = 16 bf16 per instruction)
              const uint64_t da = tc::make_desc_sw128(a0 + 32 * j);
              const uint64_t db = tc::make_desc_sw128(b0 + 32 * j);
              if (tc::elect_one()) mma_bf16(tmem_d, da, db, idesc, (c | t | j) ? 1u : 0u);
            }
          }
        }
        if (tc::elect_one()) {
          tc::umma_commit(&rg.empty[slot]);             // frees the slot once these MMAs have read it
          if (c == ngr - 1) tc::umma_commit(rg.accum);  // accumulator complete
        }
      }
      __syncwarp();
      if (++slot == (uint32_t)nslot) {
        slot = 0;
        ++use;
      }
    }
    stamp(211);
    tc::tc_fence_before();
  } else if (warp * 32 < 2 * Sg) {
    // ---------------------------- accumulator -> shared memory -------------------------------
    mbar_wait(rg.accum, ps.acc & 1);
    stamp(50);
    tc::tc_fence_after();
    const int row = warp * 32 + lane;  // TMEM lane = stacked activation row
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    float* rr = red + (size_t)row * ldred;
    for (int c = 0; c < nvalid; c += 16) {
      float a[16], b[16];
      tc::tmem_ld16x2(taddr + (uint32_t)c, taddr + (uint32_t)(nh + c), a, b);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if (c + q < nvalid) rr[c + q] = a[q] + b[q];
    }
    tc::tc_fence_before();
    stamp(51);
  }
  __syncthreads();
  tc::tc_fence_after();
  ps.cc += (uint32_t)ngr;
  ps.acc += 1;
}

// Element (stream s, column col) of an activation -> the hi / lo bf16 halves inside the tile image of chunk col / 64:
// row s (hi) and Sg + s (lo), 16-byte unit (col % 64) / 8 XOR-swizzled with the row (tc::sw128_off).
__device__ __forceinline__ void store_hl(uint8_t* img, int Sg, uint32_t tile_bytes, int s, int col, float v) {
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  const int kk = col & (KC - 1);
  uint8_t* t = img + (size_t)(col / KC) * tile_bytes + (kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(t + tc::sw128_off(s, kk >> 3)) = h;
  *reinterpret_cast<__nv_bfloat16*>(t + tc::sw128_off(Sg + s, kk >> 3)) = l;
}
__device__ __forceinline__ void tma_bulk_g2s_u32(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
}  // namespace tm

// =====================================================================================================================
// forward
// =====================================================================================================================
template <int GT>
__global__ void __launch_bounds__(kThreads, 1) lstmp_fwd_tma_kernel(const __grid_constant__ FwdTmaParams p) {
  using namespace tm;
  extern __shared__ __align__(16) uint8_t smem_raw_tma[];
  // pointer arithmetic on the __shared__ array keeps the address space visible (LDS/STS instead of generic LD/ST)
  uint8_t* base = smem_raw_tma + ((1024u - (smem_u32(smem_raw_tma) & 1023u)) & 1023u);
  uint8_t* bg = base + p.off_bg;      // gate weight slice: nch_g tiles of [roundup8(8*cpc) rows][128 B] (hi rows, then lo)
  uint8_t* bp = base + p.off_bp;      // projection slice:  nch_p tiles of [roundup8(2*rpc) rows][128 B]
  float* red = reinterpret_cast<float*>(base + p.off_red);      // [128][ldred]
  float* cprev = reinterpret_cast<float*>(base + p.off_cprev);  // [Sg*nc]
  float* peep = reinterpret_cast<float*>(base + p.off_peep);    // [3][cpc]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + p.off_bars);
  Ring rg;
  rg.ring = base + p.off_ring;
  rg.full = bars;
  rg.empty = bars + MAX_SLOTS;
  rg.accum = bars + 2 * MAX_SLOTS;
  rg.gridok = bars + 2 * MAX_SLOTS + 1;
  rg.nslot = p.nslot;
  rg.nprod = p.nprod;
  rg.gt = p.gt;
  rg.slot_bytes = p.slot_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_SLOTS + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = blockIdx.x / p.cpg, j = blockIdx.x - grp * p.cpg;  // stream group, CTA inside the group
  const int C = p.C, R = p.R, S = p.S, Sg = p.Sg, T = p.T;
  const int s_base = grp * Sg;
  const int cpc = p.cpc, rpc = p.rpc;
  const int c0 = j * cpc;
  const int nc = max(0, min(cpc, C - c0));  // my cells
  const int r0 = j * rpc;
  const int nr = max(0, min(rpc, R - r0));  // my projection outputs
  // my group's tile images in the exchange arrays [G][chunks][tile]
  const uint32_t tile_bytes = (uint32_t)(((2 * Sg + 7) & ~7) * 128);
  uint8_t* rhl = p.rhl + (size_t)grp * p.nch_g * tile_bytes;
  uint8_t* mhl = p.mhl + (size_t)grp * p.nch_p * tile_bytes;

  if (tid == 0) {
    for (int s = 0; s < p.nslot; ++s) {
      mbar_init(&rg.full[s], 1);
      mbar_init(&rg.empty[s], 1);
    }
    mbar_init(rg.accum, 1);
    mbar_init(rg.gridok, 1);
    fence_mbar_init();
  }
  if (warp == kProducerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  // zero the operand area (weight tiles + ring): rows / k tails nobody fills must not hold NaN bit patterns
  for (uint32_t o = (uint32_t)tid * 16; o < p.off_red; o += kThreads * 16)
    *reinterpret_cast<uint4*>(base + o) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  // ---- stationary weight slices: split into bf16 hi / lo once per launch ----
  if (nc > 0) {
    const int rows = 4 * nc, nk8 = R >> 3;  // hi row = gate*nc + cl (gate order g,i,f,o: LPS.h:234-243)
    for (int u = tid; u < rows * nk8; u += kThreads) {
      const int row = u % rows, k8 = u / rows;
      const int gate = row / nc, cl = row - gate * nc;
      const float4* src = reinterpret_cast<const float4*>(p.w_gifo_r + (size_t)(gate * C + c0 + cl) * R) + 2 * k8;
      const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
      const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      uint8_t* t = bg + (size_t)(k8 >> 3) * p.chunk_g;
      split8_store(v, t + tc::sw128_off(row, k8 & 7), t + tc::sw128_off(4 * cpc + row, k8 & 7));
    }
  }
  if (nr > 0) {
    const int nk8 = C >> 3;
    for (int u = tid; u < nr * nk8; u += kThreads) {
      const int row = u % nr, k8 = u / nr;
      const float4* src = reinterpret_cast<const float4*>(p.w_r_m + (size_t)(r0 + row) * C) + 2 * k8;
      const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
      const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      uint8_t* t = bp + (size_t)(k8 >> 3) * p.chunk_p;
      split8_store(v, t + tc::sw128_off(row, k8 & 7), t + tc::sw128_off(rpc + row, k8 & 7));
    }
  }

  // ---- history: c_0 of my cells -> smem and cbuf block 0; r_0 of my columns -> rbuf block 0 + hi/lo (LPS.h:231) ----
  for (int idx = tid; idx < Sg * nc; idx += kThreads) {
    int s = idx / nc, cl = idx - s * nc;
    float v = p.state_c[(size_t)(s_base + s) * C + c0 + cl];
    cprev[idx] = v;
    p.cbuf[(size_t)(s_base + s) * C + c0 + cl] = v;
  }
  for (int idx = tid; idx < Sg * nr; idx += kThreads) {
    int s = idx / nr, n = idx - s * nr;
    float v = p.state_r[(size_t)(s_base + s) * R + r0 + n];
    p.rbuf[(size_t)(s_base + s) * R + r0 + n] = v;
    store_hl(rhl, Sg, tile_bytes, s, r0 + n, v);
  }
  for (int cl = tid; cl < nc; cl += kThreads) {
    peep[cl] = p.p_i[c0 + cl];
    peep[cpc + cl] = p.p_f[c0 + cl];
    peep[2 * cpc + cl] = p.p_o[c0 + cl];
  }
  tc::fence_async_smem();  // weight tiles (generic stores) -> UMMA
  fence_async_global();    // r_0 hi/lo (generic stores) -> other CTAs' TMA
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t idesc_g = idesc_bf16(128, p.n_g), idesc_p = idesc_bf16(128, p.n_p);
  const uint32_t bg_s = smem_u32(bg), bp_s = smem_u32(bp);
  const int ldred = (int)p.ldred;
  const int rot_g = p.stagger ? (int)((unsigned)j % (unsigned)p.nch_g) : 0;
  const int rot_p = p.stagger ? (int)((unsigned)j % (unsigned)p.nch_p) : 0;

  GridSync gs;
  gs.counter = p.bar + grp * 64;
  gs.target = p.bar_base;
  gs.nctas = (unsigned)p.cpg;
  gs.off = (p.dbg & 1) != 0;
  stamp_begin((p.dbg & 4) && blockIdx.x == 0);
  __syncthreads();
  if (is_sync_thread()) gs.arrive();  // r_0 hi/lo of this CTA is in place
  stamp(1);
  Pipe ps{0u, 0u};

  for (int tt = 0; tt < T; ++tt) {
    stamp(10);
    // ================= phase 1: gates + cell update for my cells ==========================
    if (nc > 0) {
      // this thread's first element's x-part pre-activations (input GEMM + bias): in flight during the product
      float xg = 0.f, xi = 0.f, xf = 0.f, xo = 0.f;
      if (tid < Sg * nc) {
        int s = tid / nc, cl = tid - s * nc;
        const float* gp = p.gifo + (size_t)(tt * S + s_base + s) * (4 * C) + c0 + cl;
        xg = gp[0];
        xi = gp[C];
        xf = gp[2 * C];
        xo = gp[3 * C];
      }
      // gifo(t) += r(t-1) * W_gifo_r^T                                       (LPS.h:275)
      tma_product<GT>(ps, rg, gs, true, rhl, Sg, 0, p.nch_g, rot_g, bg_s, p.chunk_g, idesc_g, tmem_base + COL_A,
                  4 * cpc, 4 * nc, red, ldred);
      stamp(11);
      // pass 1: the cell update; only m(t) hi/lo -- what the other CTAs wait for -- is stored before the arrive
      for (int idx = tid; idx < Sg * nc; idx += kThreads) {
        int s = idx / nc, cl = idx - s * nc;
        size_t row = (size_t)tt * S + s_base + s;
        float* gp = p.gifo + row * (4 * C) + c0 + cl;
        if (idx >= kThreads) {
          xg = gp[0];
          xi = gp[C];
          xf = gp[2 * C];
          xo = gp[3 * C];
        }
        float* rh = red + s * ldred + cl;         // hi-activation rows
        float* rl = red + (Sg + s) * ldred + cl;  // lo-activation rows
        float cp = cprev[idx];
        float pi = peep[cl], pf = peep[cpc + cl], po = peep[2 * cpc + cl];
        float ai = (rh[nc] + rl[nc]) + xi + cp * pi;          // :278  i += c(t-1) .* peephole_i_c
        float af = (rh[2 * nc] + rl[2 * nc]) + xf + cp * pf;  // :281
        float gi = sigmoidf_fast(ai);                         // :284
        float gf = sigmoidf_fast(af);                         // :285
        float gg = tanhf_fast((rh[0] + rl[0]) + xg);          // :288
        float c = gg * gi + cp * gf;                          // :291-294
        c = fminf(fmaxf(c, -kCellClip), kCellClip);           // :296-297
        float h = tanhf_fast(c);                              // :300
        float ao = (rh[3 * nc] + rl[3 * nc]) + xo + c * po;   // :303  (uses c(t), post-clip)
        float go = sigmoidf_fast(ao);                         // :306
        float m = h * go;                                     // :309
        store_hl(mhl, Sg, tile_bytes, s, c0 + cl, m);
        cprev[idx] = c;
        // the record goes to HBM after the arrive (pass 2); park it in the consumed accumulator block meanwhile
        rh[0] = gg; rh[nc] = gi; rh[2 * nc] = gf; rh[3 * nc] = go;
        rl[0] = h; rl[nc] = m;
      }
      fence_async_global();
    } else {
      if (is_sync_thread()) gs.wait();  // no product: still wait for barrier k before arriving at k+1
    }
    __syncthreads();
    if (is_sync_thread()) gs.arrive();
    stamp(20);
    if (nc > 0) {
      // pass 2: the activation record (read by the backward launch), off the critical path
      for (int idx = tid; idx < Sg * nc; idx += kThreads) {
        int s = idx / nc, cl = idx - s * nc;
        size_t row = (size_t)tt * S + s_base + s;
        float* gp = p.gifo + row * (4 * C) + c0 + cl;
        const float* rh = red + s * ldred + cl;
        const float* rl = red + (Sg + s) * ldred + cl;
        gp[0] = rh[0];
        gp[C] = rh[nc];
        gp[2 * C] = rh[2 * nc];
        gp[3 * C] = rh[3 * nc];
        p.cbuf[(row + S) * C + c0 + cl] = cprev[idx];
        p.hbuf[row * C + c0 + cl] = rl[0];
        p.mbuf[row * C + c0 + cl] = rl[nc];
      }
      __syncthreads();  // red[] is free for the next product's read-out
    }
    // ================= phase 2: projection r(t) = m(t) * W_r_m^T for my columns (LPS.h:312) ==
    if (nr > 0) {
      tma_product<GT>(ps, rg, gs, true, mhl, Sg, 0, p.nch_p, rot_p, bp_s, p.chunk_p, idesc_p, tmem_base + COL_B,
                  rpc, nr, red, ldred);
      stamp(22);
      for (int idx = tid; idx < Sg * nr; idx += kThreads) {
        int s = idx / nr, n = idx - s * nr;
        float v = red[s * ldred + n] + red[(Sg + s) * ldred + n];
        store_hl(rhl, Sg, tile_bytes, s, r0 + n, v);
        red[s * ldred + n] = v;
      }
      fence_async_global();
    } else {
      if (is_sync_thread()) gs.wait();  // no product: still wait for barrier k before arriving at k+1
    }
    __syncthreads();
    if (tt + 1 < T && is_sync_thread()) gs.arrive();
    stamp(30);
    if (nr > 0) {
      for (int idx = tid; idx < Sg * nr; idx += kThreads) {
        int s = idx / nr, n = idx - s * nr;
        float v = red[s * ldred + n];
        size_t row = (size_t)tt * S + s_base + s;
        p.rbuf[(row + S) * R + r0 + n] = v;
        p.out[row * p.ld_out + r0 + n] = v;                               // :328
        if (tt == T - 1) p.state_r[(size_t)(s_base + s) * R + r0 + n] = v;  // :331
      }
    }
    __syncthreads();  // red[] is free for the next product's read-out
    stamp(31);
  }
  // prev_nnet_state_ <- last frame (LPS.h:331): c part
  for (int idx = tid; idx < Sg * nc; idx += kThreads) {
    int s = idx / nc, cl = idx - s * nc;
    p.state_c[(size_t)(s_base + s) * C + c0 + cl] = cprev[idx];
  }
  stamp_flush(p.dbg_stamps);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarp) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS)
                 : "memory");
  }
}

// =====================================================================================================================
// backward
// =====================================================================================================================
template <int GT>
__global__ void __launch_bounds__(kThreads, 1) lstmp_bwd_tma_kernel(const __grid_constant__ BwdTmaParams p) {
  using namespace tm;
  extern __shared__ __align__(16) uint8_t smem_raw_tma[];
  uint8_t* base = smem_raw_tma + ((1024u - (smem_u32(smem_raw_tma) & 1023u)) & 1023u);
  uint8_t* ba = base + p.off_ba;  // W_gifo_r[my K slice, my cluster's r columns]^T: tiles [roundup8(2*rpb) rows][128 B]
  uint8_t* bb = base + p.off_bb;  // W_r_m[:, my cells]^T: nch_b tiles of [roundup8(2*cpc) rows][128 B]
  float* red = reinterpret_cast<float*>(base + p.off_red);    // [roundup32(2*Sg)][ldred]  (phase A: read by the cluster)
  // Per-element state carried from step to step -- d_i(t+1), d_f(t+1), d_c(t+1) and the seven running sums of the
  // bias / peephole gradients -- lives in REGISTERS for a thread's first two elements (element idx = tid + rd * 384:
  // the same thread owns it in every timestep); only elements beyond 2 * 384 per CTA (few-CTA runs) use this buffer.
  // That frees ~17 KB of shared memory for a third ring slot at cfg3.
  float* ext = reinterpret_cast<float*>(base + p.off_ext);    // [max(0, Sg*cpc - 2*384)][10]
  float* peep = reinterpret_cast<float*>(base + p.off_peep);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + p.off_bars);
  Ring rg;
  rg.ring = base + p.off_ring;
  rg.full = bars;
  rg.empty = bars + MAX_SLOTS;
  rg.accum = bars + 2 * MAX_SLOTS;
  rg.gridok = bars + 2 * MAX_SLOTS + 1;
  rg.nslot = p.nslot;
  rg.nprod = p.nprod;
  rg.gt = p.gt;
  rg.slot_bytes = p.slot_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_SLOTS + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = blockIdx.x / p.cpg, j = blockIdx.x - grp * p.cpg;  // stream group, CTA inside the group
  const int C = p.C, R = p.R, S = p.S, Sg = p.Sg, T = p.T;
  const int s_base = grp * Sg;
  const int cpc = p.cpc, rpb = p.rpb, kp = p.kp;
  const int a = (int)cluster_ctarank();  // K slice of the d_r product
  const int b = j / kp;                  // cluster index inside the group: r columns [n0, n0 + nn)
  const int n0 = b * rpb;
  const int nn = max(0, min(rpb, R - n0));
  const int c0 = j * cpc;
  const int nc = max(0, min(cpc, C - c0));  // my cells
  const uint32_t tile_bytes = (uint32_t)(((2 * Sg + 7) & ~7) * 128);
  uint8_t* dghl = p.dghl + (size_t)grp * p.nch_a * tile_bytes;
  uint8_t* drhl = p.drhl + (size_t)grp * p.nch_b * tile_bytes;
  // my K slice of the 4C contraction, in 64-column chunks
  const int kbase = p.nch_a / kp, krem = p.nch_a % kp;
  const int ks = a * kbase + min(a, krem);
  const int nka = kbase + (a < krem ? 1 : 0);

  if (tid == 0) {
    for (int s = 0; s < p.nslot; ++s) {
      mbar_init(&rg.full[s], 1);
      mbar_init(&rg.empty[s], 1);
    }
    mbar_init(rg.accum, 1);
    mbar_init(rg.gridok, 1);
    fence_mbar_init();
  }
  if (warp == kProducerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  for (uint32_t o = (uint32_t)tid * 16; o < p.off_red; o += kThreads * 16)
    *reinterpret_cast<uint4*>(base + o) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  // ---- stationary weight slices, transposed on the way in (B[n][k] = W[k][n]) and split into bf16 hi / lo ----
  if (nn > 0) {
    const int K4 = 4 * C;
    for (int u = tid; u < nn * nka * 8; u += kThreads) {
      const int n = u % nn, k8 = u / nn;  // k8: 8-k unit inside my slice
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = ks * KC + 8 * k8 + q;
        v[q] = k < K4 ? __ldg(p.w_gifo_r + (size_t)k * R + n0 + n) : 0.f;
      }
      uint8_t* t = ba + (size_t)(k8 >> 3) * p.chunk_a;
      split8_store(v, t + tc::sw128_off(n, k8 & 7), t + tc::sw128_off(rpb + n, k8 & 7));
    }
  }
  if (nc > 0) {
    for (int u = tid; u < nc * p.nch_b * 8; u += kThreads) {
      const int n = u % nc, k8 = u / nc;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = 8 * k8 + q;
        v[q] = k < R ? __ldg(p.w_r_m + (size_t)k * C + c0 + n) : 0.f;
      }
      uint8_t* t = bb + (size_t)(k8 >> 3) * p.chunk_b;
      split8_store(v, t + tc::sw128_off(n, k8 & 7), t + tc::sw128_off(cpc + n, k8 & 7));
    }
  }
  float st0[10], st1[10];  // [0] d_i(t+1), [1] d_f(t+1), [2] d_c(t+1) (row-block T+1 is zero, LPS.h:352), [3..9] sums
#pragma unroll
  for (int q = 0; q < 10; ++q) st0[q] = st1[q] = 0.f;
  for (int idx = tid; idx < (Sg * cpc - 2 * kThreads) * 10; idx += kThreads) ext[idx] = 0.f;
  for (int cl = tid; cl < nc; cl += kThreads) {
    peep[cl] = p.p_i[c0 + cl];
    peep[cpc + cl] = p.p_f[c0 + cl];
    peep[2 * cpc + cl] = p.p_o[c0 + cl];
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t idesc_a = idesc_bf16(128, p.n_a), idesc_b = idesc_bf16(128, p.n_b);
  const uint32_t ba_s = smem_u32(ba), bb_s = smem_u32(bb);
  const int ldred = (int)p.ldred;
  const int rot_a = (p.stagger && nka > 0) ? (int)((unsigned)b % (unsigned)nka) : 0;
  const int rot_b = p.stagger ? (int)((unsigned)j % (unsigned)p.nch_b) : 0;
  const uint32_t red_s = smem_u32(red);

  GridSync gs;
  gs.counter = p.bar + grp * 64;
  gs.target = p.bar_base;
  gs.nctas = (unsigned)p.cpg;
  gs.off = (p.dbg & 1) != 0;
  stamp_begin((p.dbg & 8) && blockIdx.x == 0);
  cluster_sync_all();  // every CTA of the cluster is running (its shared memory may be read from now on)
  stamp(2);
  Pipe ps{0u, 0u};
  // my share of the cluster's [Sg x nn] d_r block in the reduction: streams s = a (mod kp)
  const int nsh = (Sg - a + kp - 1) / kp;  // streams a, a+kp, ...

  for (int tt = T - 1; tt >= 0; --tt) {
    const bool have_next = (tt + 1 < T);
    stamp(40);
    // ============ phase A: d_r(t)[:, my columns] = out_diff(t) + DGIFO(t+1) * W_gifo_r            (LPS.h:367,391)
    if (nn > 0) {
      // out_diff of this thread's first reduction element: in flight during the product
      float od = 0.f;
      if (tid < nsh * nn) {
        const int s = a + kp * (tid / nn), n = tid % nn;
        od = p.out_diff[((size_t)tt * S + s_base + s) * p.ld_od + n0 + n];
      }
      if (have_next) {
        if (nka > 0) {
          tma_product<GT>(ps, rg, gs, true, dghl, Sg, ks, nka, rot_a, ba_s, p.chunk_a, idesc_a,
                      tmem_base + COL_A, rpb, nn, red, ldred);
        } else {
          if (is_sync_thread()) gs.wait();
          for (int idx = tid; idx < ((2 * Sg + 31) & ~31) * ldred; idx += kThreads) red[idx] = 0.f;
        }
        stamp(41);
        cluster_sync_all();  // the kp partial blocks of this cluster are complete
        stamp(42);
      }
      for (int li = tid; li < nsh * nn; li += kThreads) {
        const int s = a + kp * (li / nn), n = li % nn;
        const size_t row = (size_t)tt * S + s_base + s;
        float v = (li < kThreads) ? od : p.out_diff[row * p.ld_od + n0 + n];
        if (have_next) {
          // the kp partial blocks through distributed shared memory, all loads in flight, fixed summation order
          const uint32_t off_h = red_s + (uint32_t)((s * ldred + n) * 4);
          const uint32_t off_l = red_s + (uint32_t)(((Sg + s) * ldred + n) * 4);
          float ph[8], pl[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            ph[q] = q < kp ? ld_dsmem_f32(off_h, (uint32_t)q) : 0.f;
            pl[q] = q < kp ? ld_dsmem_f32(off_l, (uint32_t)q) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) v += ph[q] + pl[q];
        }
        store_hl(drhl, Sg, tile_bytes, s, n0 + n, v);
        p.dr[row * R + n0 + n] = v;
      }
      fence_async_global();
    } else if (have_next) {
      if (is_sync_thread()) gs.wait();
    }
    __syncthreads();
    if (is_sync_thread()) gs.arrive();
    stamp(43);
    // ============ phase B: d_m = d_r * W_r_m (my cells) (:408) + gate derivatives (:411-440)
    if (nc > 0) {
      // prefetch this thread's first TWO elements' activations while d_r is gathered and contracted (with two stream
      // groups and 4-CTA clusters a CTA owns 13 cells x 32 streams = 416 elements: a second round for 32 threads)
      float yv[2][8];
#pragma unroll
      for (int rd = 0; rd < 2; ++rd) {
        const int idx = tid + rd * kThreads;
#pragma unroll
        for (int q = 0; q < 8; ++q) yv[rd][q] = 0.f;
        if (idx < Sg * nc) {
          int s = idx / nc, cl = idx - s * nc;
          size_t row = (size_t)tt * S + s_base + s;
          const float* gp = p.gifo + row * (4 * C) + c0 + cl;
          yv[rd][0] = gp[0]; yv[rd][1] = gp[C]; yv[rd][2] = gp[2 * C]; yv[rd][3] = gp[3 * C];
          yv[rd][4] = p.cbuf[(row + S) * C + c0 + cl];
          yv[rd][5] = p.cbuf[row * C + c0 + cl];
          yv[rd][6] = p.hbuf[row * C + c0 + cl];
          yv[rd][7] = have_next ? gp[(size_t)S * 4 * C + 2 * C] : 0.f;
        }
      }
      tma_product<GT>(ps, rg, gs, true, drhl, Sg, 0, p.nch_b, rot_b, bb_s, p.chunk_b, idesc_b,
                  tmem_base + COL_B, cpc, nc, red, ldred);
      stamp(45);
      const int n_el = Sg * nc;
      // one element (stream s, cell cl): y = its activation record, st = its carried state
      auto element = [&](int idx, const float (&y)[8], float (&st)[10]) {
        const int s = idx / nc, cl = idx - s * nc;
        const float yg = y[0], yi = y[1], yf = y[2], yo = y[3], yc = y[4], ycp = y[5], yh = y[6], yfn = y[7];
        const float pi = peep[cl], pf = peep[cpc + cl], po = peep[2 * cpc + cl];
        const float d_m = red[s * ldred + cl] + red[(Sg + s) * ldred + cl];
        const float d_h = (d_m * yo) * (1.0f - yh * yh);      // :411-412
        const float d_o = (d_m * yh) * yo * (1.0f - yo);      // :415-416
        float d_c = d_h;                                      // :424
        d_c += st[2] * yfn;                                   // :425
        d_c += st[0] * pi;                                    // :426
        d_c += st[1] * pf;                                    // :427
        d_c += d_o * po;                                      // :428
        const float d_f = (d_c * ycp) * yf * (1.0f - yf);     // :431-432
        const float d_i = (d_c * yg) * yi * (1.0f - yi);      // :435-436
        const float d_g = (d_c * yi) * (1.0f - yg * yg);      // :439-440
        // DGIFO(t) hi/lo is what the other CTAs wait for; the fp32 record is stored after the arrive
        store_hl(dghl, Sg, tile_bytes, s, c0 + cl, d_g);
        store_hl(dghl, Sg, tile_bytes, s, C + c0 + cl, d_i);
        store_hl(dghl, Sg, tile_bytes, s, 2 * C + c0 + cl, d_f);
        store_hl(dghl, Sg, tile_bytes, s, 3 * C + c0 + cl, d_o);
        st[0] = d_i;
        st[1] = d_f;
        st[2] = d_c;
        red[s * ldred + cl] = d_g;         // parked for pass 2 (d_i, d_f are in the carried state)
        red[(Sg + s) * ldred + cl] = d_o;
        st[3] += d_g;          // bias_corr_ column sums (:474)
        st[4] += d_i;
        st[5] += d_f;
        st[6] += d_o;
        st[7] += d_i * ycp;    // peephole_i_c_corr_  DI(t) .* C(t-1)  (:477)
        st[8] += d_f * ycp;    // peephole_f_c_corr_                  (:480)
        st[9] += d_o * yc;     // peephole_o_c_corr_  DO(t) .* C(t)    (:483)
      };
      if (tid < n_el) element(tid, yv[0], st0);
      if (tid + kThreads < n_el) element(tid + kThreads, yv[1], st1);
      for (int idx = tid + 2 * kThreads; idx < n_el; idx += kThreads) {
        const int s = idx / nc, cl = idx - s * nc;
        const size_t row = (size_t)tt * S + s_base + s;
        const float* gp = p.gifo + row * (4 * C) + c0 + cl;
        const float y[8] = {gp[0], gp[C], gp[2 * C], gp[3 * C], p.cbuf[(row + S) * C + c0 + cl],
                            p.cbuf[row * C + c0 + cl] /* c(t-1): block tt */, p.hbuf[row * C + c0 + cl],
                            have_next ? gp[(size_t)S * 4 * C + 2 * C] : 0.f /* f(t+1) */};
        float* e = ext + (size_t)(idx - 2 * kThreads) * 10;
        float st[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) st[q] = e[q];
        element(idx, y, st);
#pragma unroll
        for (int q = 0; q < 10; ++q) e[q] = st[q];
      }
      fence_async_global();
    } else {
      if (is_sync_thread()) gs.wait();
    }
    __syncthreads();
    if (tt > 0 && is_sync_thread()) gs.arrive();
    stamp(46);
    if (nc > 0) {
      for (int idx = tid; idx < Sg * nc; idx += kThreads) {
        int s = idx / nc, cl = idx - s * nc;
        size_t row = (size_t)tt * S + s_base + s;
        float* dp = p.dgifo + row * (4 * C) + c0 + cl;
        const float* e = ext + (size_t)(idx - 2 * kThreads) * 10;  // (only dereferenced for idx >= 2 * 384)
        dp[0] = red[s * ldred + cl];
        dp[C] = idx < kThreads ? st0[0] : idx < 2 * kThreads ? st1[0] : e[0];
        dp[2 * C] = idx < kThreads ? st0[1] : idx < 2 * kThreads ? st1[1] : e[1];
        dp[3 * C] = red[(Sg + s) * ldred + cl];
      }
    }
    __syncthreads();  // red[] is free for the next product's read-out
    stamp(47);
  }
  // bias / peephole gradients of my cells: sum over the streams in a fixed order; per-group partials when the streams
  // are split into groups (summed by small_grads_kernel), else straight into the gradient arena
  // (bias(4C) | peephole_i | peephole_f | peephole_o are contiguous there, LPS.h:162-189)
  __syncthreads();
  float* acc7 = reinterpret_cast<float*>(rg.ring);  // [Sg*nc][7]: the operand ring is idle now
  for (int idx = tid; idx < Sg * nc; idx += kThreads) {
    const float* e = ext + (size_t)(idx - 2 * kThreads) * 10;
#pragma unroll
    for (int w = 0; w < 7; ++w)
      acc7[(size_t)idx * 7 + w] = idx < kThreads ? st0[3 + w] : idx < 2 * kThreads ? st1[3 + w] : e[3 + w];
  }
  __syncthreads();
  float* gsm = p.g_small + (size_t)grp * 7 * C;
  for (int q = tid; q < nc * 7; q += kThreads) {
    int cl = q / 7, w = q - cl * 7;
    float s7 = 0.f;
    for (int s = 0; s < Sg; ++s) s7 += acc7[(size_t)(s * nc + cl) * 7 + w];
    gsm[(size_t)w * C + c0 + cl] = s7;
  }
  stamp_flush(p.dbg_stamps);
  tc::tc_fence_before();
  cluster_sync_all();  // nobody exits while a peer may still read its partial block
  if (warp == kProducerWarp) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static size_t round1k(size_t b) { return (b + 1023) & ~size_t(1023); }
static int max_slots();
static int num_producers() {  // LSTMP_B200_TMA_PRODUCERS: 1..3 producer warps (default 3)
  const char* v = getenv("LSTMP_B200_TMA_PRODUCERS");
  const int n = (v && *v) ? atoi(v) : tm::NPROD;
  return n < 1 ? 1 : n > tm::NPROD ? tm::NPROD : n;
}
// 64-k tiles per bulk copy / ring slot: 4, 2 or 1.  Forward default 4 (LSTMP_B200_TMA_GROUP_TILES), backward default 2
// (LSTMP_B200_TMA_GROUP_TILES_BWD; its shared memory holds two weight slices and leaves room for a short ring only).
// Measured at cfg3 (us per launch, forward / backward): 1 tile 208 / 290, 2 tiles 191 / 271, 4 tiles 182 / 278 (G = 1).
static int group_tiles(bool bwd) {
  const char* v = getenv(bwd ? "LSTMP_B200_TMA_GROUP_TILES_BWD" : "LSTMP_B200_TMA_GROUP_TILES");
  const int n = (v && *v) ? atoi(v) : (bwd ? 2 : 4);
  return n >= 4 ? 4 : n >= 2 ? 2 : 1;
}
// ring geometry: the most tiles per slot (<= the wish) that still leaves two slots in `avail` bytes
static bool ring_plan(size_t avail, int Sg, bool bwd, int* gt, unsigned* slot_bytes, int* nslot) {
  const size_t tile = (size_t)((2 * Sg + 7) & ~7) * 128;
  for (int g = group_tiles(bwd); g >= 1; g >>= 1) {
    const size_t sb = g * tile;
    // (an MMA reads 128 rows of a tile whatever Sg is: a slot is rounded up so that the over-read stays inside it)
    const size_t need = sb + (tile < tm::TILE_MAX ? tm::TILE_MAX - tile : 0);
    if (avail >= 2 * need) {
      int n = (int)(avail / need);
      if (n > max_slots()) n = max_slots();
      *gt = g;
      *slot_bytes = (unsigned)((need + 1023) & ~size_t(1023));
      *nslot = n;
      return true;
    }
  }
  return false;
}
static int max_slots() {  // LSTMP_B200_TMA_SLOTS caps the ring depth (tests: the 2-slot ring of the tightest shapes)
  const char* v = getenv("LSTMP_B200_TMA_SLOTS");
  const int n = (v && *v) ? atoi(v) : tm::MAX_SLOTS;
  return n < 2 ? 2 : n > tm::MAX_SLOTS ? tm::MAX_SLOTS : n;
}

bool fwd_tma_plan(int C, int R, int S, int G, int nctas, size_t smem_limit, FwdTmaParams* p, size_t* smem_bytes) {
  using namespace tm;
  if (G < 1 || S % G || nctas < G) return false;
  const int Sg = S / G, cpg = nctas / G;
  if (Sg < 1 || Sg > 64 || (C & 7) || (R & 7)) return false;
  const int cpc = (C + cpg - 1) / cpg, rpc = (R + cpg - 1) / cpg;
  const int n_g = (8 * cpc + 15) & ~15, n_p = (2 * rpc + 15) & ~15;
  if (n_g > 256 || n_p > 256) return false;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += round1k(bytes);
    return (unsigned)o;
  };
  p->G = G;
  p->Sg = Sg;
  p->cpg = cpg;
  p->cpc = cpc;
  p->rpc = rpc;
  p->n_g = n_g;
  p->n_p = n_p;
  p->nctas = cpg * G;
  p->nch_g = (R + KC - 1) / KC;
  p->nch_p = (C + KC - 1) / KC;
  p->chunk_g = (unsigned)(((8 * cpc + 7) & ~7) * 128);
  p->chunk_p = (unsigned)(((2 * rpc + 7) & ~7) * 128);
  p->off_bg = take((size_t)p->nch_g * p->chunk_g);
  p->off_bp = take((size_t)p->nch_p * p->chunk_p);
  // an MMA reads N (>= actual) weight rows per tile: the over-read of the last tiles lands in the ring that follows
  // (finite values only -- the area is zeroed at kernel start -- and only into accumulator columns nobody reads)
  p->off_ring = (unsigned)off;
  const int ldred = ((4 * cpc > rpc ? 4 * cpc : rpc) | 1);
  p->ldred = (unsigned)ldred;
  const size_t tail = round1k((size_t)128 * ldred * 4) + round1k((size_t)Sg * cpc * 4) + 1024 /* peepholes */ +
                      1024 /* barriers + tmem slot */;
  const size_t reserve = 1024 /* base alignment */ + (size_t)static_smem_reserve();
  if (smem_limit < off + tail + reserve) return false;
  int nslot = 0, gt = 1;
  unsigned slot_bytes = 0;
  if (!ring_plan(smem_limit - off - tail - reserve, Sg, false, &gt, &slot_bytes, &nslot)) return false;
  p->nslot = nslot;
  p->nprod = num_producers();
  p->gt = gt;
  p->slot_bytes = slot_bytes;
  off += (size_t)nslot * slot_bytes;
  p->off_red = take((size_t)128 * ldred * 4);
  p->off_cprev = take((size_t)Sg * cpc * 4);
  p->off_peep = take((size_t)3 * cpc * 4);
  p->off_bars = take(256);
  *smem_bytes = off + 1024;
  return *smem_bytes + (size_t)static_smem_reserve() <= smem_limit;
}

// nctas: CTAs of the whole grid (G groups of nctas / G, each a whole number of kp-CTA clusters)
bool bwd_tma_plan(int C, int R, int S, int G, int nctas, int kp, size_t smem_limit, BwdTmaParams* p, size_t* smem_bytes) {
  using namespace tm;
  if (G < 1 || S % G || kp < 1 || kp > 8) return false;
  const int Sg = S / G, cpg = nctas / G / kp * kp;
  if (cpg < kp || Sg < 1 || Sg > 64 || (C & 7) || (R & 7)) return false;
  const int np = cpg / kp;
  const int cpc = (C + cpg - 1) / cpg, rpb = (R + np - 1) / np;
  const int n_a = (2 * rpb + 15) & ~15, n_b = (2 * cpc + 15) & ~15;
  if (n_a > 256 || n_b > 256) return false;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += round1k(bytes);
    return (unsigned)o;
  };
  p->G = G;
  p->Sg = Sg;
  p->cpg = cpg;
  p->nctas = cpg * G;
  p->kp = kp;
  p->cpc = cpc;
  p->rpb = rpb;
  p->n_a = n_a;
  p->n_b = n_b;
  p->nch_a = (4 * C + KC - 1) / KC;
  p->nch_b = (R + KC - 1) / KC;
  const int nka_max = (p->nch_a + kp - 1) / kp;
  p->chunk_a = (unsigned)(((2 * rpb + 7) & ~7) * 128);
  p->chunk_b = (unsigned)(((2 * cpc + 7) & ~7) * 128);
  p->off_ba = take((size_t)nka_max * p->chunk_a);
  p->off_bb = take((size_t)p->nch_b * p->chunk_b);
  p->off_ring = (unsigned)off;
  const int ldred = ((rpb > cpc ? rpb : cpc) | 1);
  p->ldred = (unsigned)ldred;
  const int red_rows = (2 * Sg + 31) & ~31;
  const size_t n_ext = Sg * cpc > 2 * kThreads ? (size_t)(Sg * cpc - 2 * kThreads) : 0;
  // (the bias / peephole sums are reduced through the idle ring at the end of the kernel: [Sg*cpc][7] floats)
  const size_t tail = round1k((size_t)red_rows * ldred * 4) + round1k(n_ext * 10 * 4) + 1024 + 1024;
  const size_t reserve = 1024 + (size_t)static_smem_reserve();
  if (smem_limit < off + tail + reserve) return false;
  int nslot = 0, gt = 1;
  unsigned slot_bytes = 0;
  if (!ring_plan(smem_limit - off - tail - reserve, Sg, true, &gt, &slot_bytes, &nslot)) return false;
  p->nslot = nslot;
  p->nprod = num_producers();
  p->gt = gt;
  p->slot_bytes = slot_bytes;
  off += (size_t)nslot * slot_bytes;
  if ((size_t)nslot * slot_bytes < (size_t)Sg * cpc * 7 * 4) return false;
  p->off_red = take((size_t)red_rows * ldred * 4);
  p->off_ext = take(n_ext * 10 * 4);
  p->off_peep = take((size_t)3 * cpc * 4);
  p->off_bars = take(256);
  *smem_bytes = off + 1024;
  return *smem_bytes + (size_t)static_smem_reserve() <= smem_limit;
}

cudaError_t tma_set_smem_limits(size_t fwd_bytes, size_t bwd_bytes) {
  // per device and per function; several engines of different shapes may share the process: only ever raise it
  static size_t cur_f[64] = {0}, cur_b[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  dev &= 63;
  if (fwd_bytes > cur_f[dev]) {
    for (const void* fn : {(const void*)lstmp_fwd_tma_kernel<1>, (const void*)lstmp_fwd_tma_kernel<2>,
                           (const void*)lstmp_fwd_tma_kernel<4>}) {
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_bytes);
      if (e != cudaSuccess) return e;
    }
    cur_f[dev] = fwd_bytes;
  }
  if (bwd_bytes > cur_b[dev]) {
    for (const void* fn : {(const void*)lstmp_bwd_tma_kernel<1>, (const void*)lstmp_bwd_tma_kernel<2>,
                           (const void*)lstmp_bwd_tma_kernel<4>}) {
      e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_bytes);
      if (e != cudaSuccess) return e;
    }
    cur_b[dev] = bwd_bytes;
  }
  return cudaSuccess;
}

// Largest grid (multiple of kp, <= max_ctas) of kp-CTA clusters that is co-resident on the device.
int bwd_tma_max_ctas(int kp, size_t smem_bytes, int max_ctas) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(max_ctas / kp * kp));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)kp;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, (const void*)lstmp_bwd_tma_kernel<1>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int n = ncl * kp;
  if (n > max_ctas) n = max_ctas / kp * kp;
  return n;
}

cudaError_t launch_fwd_tma(const FwdTmaParams& p, size_t smem_bytes, cudaStream_t stream) {
  void* args[] = {(void*)&p};
  dim3 grid(p.nctas), block(kThreads);
  const void* fn = p.gt == 4 ? (const void*)lstmp_fwd_tma_kernel<4>
                   : p.gt == 2 ? (const void*)lstmp_fwd_tma_kernel<2> : (const void*)lstmp_fwd_tma_kernel<1>;
  return cudaLaunchCooperativeKernel(fn, grid, block, args, smem_bytes, stream);
}

cudaError_t launch_bwd_tma(const BwdTmaParams& p, size_t smem_bytes, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.nctas);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  void (*fn)(const BwdTmaParams) = p.gt == 4 ? lstmp_bwd_tma_kernel<4> : p.gt == 2 ? lstmp_bwd_tma_kernel<2>
                                                                                   : lstmp_bwd_tma_kernel<1>;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p.kp;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  // co-residency of the whole grid is what the grid barrier needs: the grid was sized with
  // cudaOccupancyMaxActiveClusters; the cooperative attribute makes the driver check it too.  Drivers that refuse
  // cluster + cooperative in one launch get the plain cluster launch (same co-residency by construction).
  // LSTMP_B200_BWD_COOP=0 skips the cooperative attribute (Nsight Compute cannot replay a cooperative cluster launch)
  static int coop_ok = [] {
    const char* v = getenv("LSTMP_B200_BWD_COOP");
    return (v && *v) ? atoi(v) : 1;
  }();
  if (coop_ok) {
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, p);
    if (e == cudaSuccess) return e;
    cudaGetLastError();
    coop_ok = 0;
  }
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, fn, p);
}

}  // namespace lstmp
