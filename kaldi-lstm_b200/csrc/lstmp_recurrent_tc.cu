// Tensor-core (tcgen05) forward time loop for LstmProjectedStreams on sm_100a.
//
// Same contract, record layout and results as lstmp_fwd_kernel (lstmp_recurrent.cu; reference
// google/nnet/bd-nnet-lstm-projected-streams.h:261-331), for num_stream <= 64: ONE cooperative launch per chunk,
// one CTA per SM, every CTA keeps its slice of W_gifo_r / W_r_m in shared memory for the whole chunk.  The two
// per-timestep contractions run on the 5th-gen tensor cores with FP32-faithful split arithmetic:
//
//   D[128 x N] (TMEM, fp32) = A[128 x K] * B[N x K]^T        tcgen05.mma.kind::tf32, M = 128, K = 8 per instruction
//
//   A = the all-gathered activations of ALL streams (r(t-1) for the gates, m(t) for the projection), rows 0..S-1
//       the "hi" halves (x & 0xffffe000, exact TF32) and rows S..2S-1 the "lo" halves (x - hi), written by 8
//       loader warps (cp.async.cg from L2 -> landing slot -> split -> st.shared in the SWIZZLE_128B K-major layout)
//       into a ring of [128 rows x 32 k] slots guarded by full/ready/empty mbarriers;
//   B = the CTA's stationary weight slice, hi rows then lo rows, split ONCE per launch.
//
// One MMA therefore yields all four hi/lo cross products: result[s][n] = D[s][n] + D[s][nh+n] + D[S+s][n] +
// D[S+s][nh+n] (the lo*lo term costs nothing extra).  Per MMA the tensor pipe needs max(N/2, ~40) cycles (N = 48 for
// the gates of 6 cells, 16 for 4 projection columns).  Measured: 367 us per launch vs 423 us for the
// FFMA kernel; a chunk of 32 k still takes ~600 cycles, far above what L2 (55-72 B/clk/SM) and the MMAs (180 cycles)
// need -- see the experiment log for what was ruled out.
#include "lstmp_common.cuh"
#include "lstmp_kernels.h"
#include "lstmp_tc.cuh"

namespace lstmp {

namespace tcf {
constexpr int KC = 32;                          // k per ring slot (4 MMAs of K = 8)
constexpr int SLABS = KC / 4;                   // 16-byte K slabs per slot
constexpr int LOADERS = 256;                    // warps 0-7
constexpr int PF = 3;                           // LOADER 1: chunks of cp.async copies each loader thread keeps in flight
constexpr int PFR = 4;                          // LOADER 0: chunks of LDG.128 (register sets) in flight per loader thread
constexpr uint32_t SLOT_BYTES = 128 * 128;      // activation slot: 128 rows x 128 B, SWIZZLE_128B (lstmp_tc.cuh)
constexpr int MAX_SLOTS = 8;
constexpr uint32_t TMEM_COLS = 512;             // the whole TMEM: this CTA is alone on its SM
constexpr uint32_t COL_G = 0, COL_P = 256;      // accumulator columns of the gate / projection product

struct Pipe {
  uint32_t cc;   // ring chunks produced/consumed so far (same sequence in every thread)
  uint32_t acc;  // products finished so far (accumulator-ready phase)
};

__device__ __forceinline__ void split4(float4 x, float4& h, float4& l) {
  h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
  h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
  h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
  h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
  l.x = x.x - h.x;
  l.y = x.y - h.y;
  l.z = x.z - h.z;
  l.w = x.w - h.w;
}

// red[row*ldred + n] = D[row][n] + D[row][nh + n]  for the 128 stacked rows, n < nvalid.
// X: [S x K] activations in global memory (row stride ld), written by other CTAs before the preceding group barrier.
// Every thread of the CTA calls this; contains one __syncthreads.
template <int LOADER>  // 0: LDG.128 register prefetch, 1: cp.async landing slots (default), 2: warp-per-chunk
__device__ __forceinline__ void tc_product(const FwdTcParams& p, Pipe& ps, const float* __restrict__ X, int ld, int K,
                                           uint32_t b_addr, uint32_t chunk_b, uint32_t idesc, uint32_t tmem_d, int nh,
                                           int nvalid, uint8_t* ring, uint64_t* full, uint64_t* ready,
                                           uint64_t* empty, uint64_t* accum, float* red, uint8_t* stage) {
  // warp index through a shuffle: provably warp-uniform for ptxas, so everything computed inside the issuer branch
  // stays in uniform registers (no R2UR per MMA)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int nch = K / KC;
  const int nslot = p.nslot;
  uint32_t slot = ps.cc % (uint32_t)nslot, use = ps.cc / (uint32_t)nslot;
  // Every CTA of the group gathers the same [S x K] block.  CTA j walks the K chunks starting at chunk (j mod nch)
  // so that at any moment the SMs ask L2 for different lines (the sum over k is order-independent up to fp32
  // rounding, and the order is fixed per CTA: results stay bit-reproducible).
  const int rot = p.stagger ? (int)(blockIdx.x % (unsigned)nch) : 0;

  if (warp < 8) {
    // ------------------------------ loader / transform ---------------------------------------
    // LOADER 1 (default): stage 1 = cp.async.cg (LDGSTS, L2 -> raw landing slots, PF chunks in flight per thread,
    // tracked by cp.async groups); stage 2 = each thread reads back ITS OWN two 16-byte units, splits them and stores
    // hi / lo into the UMMA tile.  The register-prefetch variant below measured 10 % slower (405 vs 367 us).
    const int S = p.S;
    const int nunits = S * SLABS;  // float4 units per chunk (<= 512)
    const int u0 = tid, u1 = tid + LOADERS;
    const bool ok0 = u0 < nunits, ok1 = u1 < nunits;
    const int r0 = u0 >> 3, r1 = u1 >> 3, kc = tid & 7;
    const float* g0 = X + (size_t)r0 * ld + 4 * kc;
    const float* g1 = X + (size_t)r1 * ld + 4 * kc;
    // hi row r -> tile row r, lo row r -> tile row S + r (S % 8 == 0: same swizzle phase, offset S*128 bytes)
    const uint32_t so0 = tc::sw128_off(r0, kc), so1 = tc::sw128_off(r1, kc), lo_off = (uint32_t)S * 128;
    if constexpr (LOADER == 1) {
      uint8_t* my0 = stage + (size_t)u0 * 16;
      uint8_t* my1 = stage + (size_t)u1 * 16;
      const size_t stage_bytes = (size_t)nunits * 16;
      auto issue = [&](int c, int st) {
        const int ce = (c + rot < nch) ? c + rot : c + rot - nch;
        if (ok0) cp_async16(my0 + st * stage_bytes, g0 + ce * KC);
        if (ok1) cp_async16(my1 + st * stage_bytes, g1 + ce * KC);
      };
  #pragma unroll
      for (int i = 0; i < PF; ++i) {
        if (i < nch) issue(i, i);
        cp_async_commit();
      }
      for (int c0 = 0; c0 < nch; c0 += PF) {
  #pragma unroll
        for (int i = 0; i < PF; ++i) {
          const int c = c0 + i;
          if (c < nch) {
            cp_async_wait<PF - 1>();  // this thread's copies of chunk c have landed
          stamp(50);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 x0 = ok0 ? *reinterpret_cast<const float4*>(my0 + i * stage_bytes) : z;
            const float4 x1 = ok1 ? *reinterpret_cast<const float4*>(my1 + i * stage_bytes) : z;
            if (c + PF < nch) issue(c + PF, i);
            cp_async_commit();  // one group per iteration (possibly empty) keeps the wait_group arithmetic uniform
            if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
            uint8_t* st = ring + (size_t)slot * SLOT_BYTES;
            float4 h, l;
            if (ok0) {
              split4(x0, h, l);
              *reinterpret_cast<float4*>(st + so0) = h;
              *reinterpret_cast<float4*>(st + so0 + lo_off) = l;
            }
            if (ok1) {
              split4(x1, h, l);
              *reinterpret_cast<float4*>(st + so1) = h;
              *reinterpret_cast<float4*>(st + so1 + lo_off) = l;
            }
            // The proxy fence is on the consumer side (issuer warp, after its acquire-wait): fence.proxy.async is
            // MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in SASS.  The release-arrive orders the stores.
            // ONE arrival per warp: 256 per-thread arrivals on the same mbarrier serialise (~3 cycles each) and were the
            // whole cost of a chunk (~800 cycles).  __syncwarp orders the other lanes' stores before lane 0's release.
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[slot]);
            if (++slot == (uint32_t)nslot) {
              slot = 0;
              ++use;
            }
          }
        }
      }
    } else if constexpr (LOADER == 2) {
      // Warp-per-chunk (LSTMP_B200_TC_LOADER=2, not yet validated on hardware): warp w loads, splits and stores the
      // chunks c = w (mod 8) on its own -- lane l owns the 16-byte column l & 7 of rows (l >> 3) + 4 i -- and lane 0
      // arrives on a full barrier of count 1, so up to nslot chunks are in progress at once.  In the two lockstep
      // layouts every warp takes part in every chunk and a chunk cannot complete faster than one warp's dependent
      // chain wait -> read back -> split -> wait for the slot -> store -> arrive.
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kcw = lane & 7, rl = lane >> 3;
      for (int c = 0; c < nch; ++c) {
        if ((c & 7) == warp) {
          const int ce = (c + rot < nch) ? c + rot : c + rot - nch;
          const float* gsrc = X + (size_t)rl * ld + ce * KC + 4 * kcw;
          float4 x[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = (rl + 4 * i < S) ? ld_cg_f4(gsrc + (size_t)(4 * i) * ld) : z;
          if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
          uint8_t* st = ring + (size_t)slot * SLOT_BYTES;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = rl + 4 * i;
            if (r < S) {
              float4 h, l;
              split4(x[i], h, l);
              const uint32_t so = tc::sw128_off(r, kcw);
              *reinterpret_cast<float4*>(st + so) = h;
              *reinterpret_cast<float4*>(st + so + lo_off) = l;
            }
          }
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&full[slot]);
        }
        if (++slot == (uint32_t)nslot) {
          slot = 0;
          ++use;
        }
      }
    } else {
      // Register prefetch (LSTMP_B200_TC_LOADER=0): PFR chunks of LDG.128 (ld.global.cg) in flight per thread, no
      // landing slots -- a third less shared-memory traffic per chunk and room for a 6-slot ring, yet slower.
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v[PFR][2];
#pragma unroll
      for (int i = 0; i < PFR; ++i) {
        v[i][0] = z;
        v[i][1] = z;
        if (i < nch) {
          const int ce = (i + rot < nch) ? i + rot : i + rot - nch;
          if (ok0) v[i][0] = ld_cg_f4(g0 + ce * KC);
          if (ok1) v[i][1] = ld_cg_f4(g1 + ce * KC);
        }
      }
      for (int c0 = 0; c0 < nch; c0 += PFR) {
#pragma unroll
        for (int i = 0; i < PFR; ++i) {
          const int c = c0 + i;
          if (c < nch) {
            if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
            uint8_t* st = ring + (size_t)slot * SLOT_BYTES;
            float4 h, l;
            if (ok0) {
              split4(v[i][0], h, l);
              *reinterpret_cast<float4*>(st + so0) = h;
              *reinterpret_cast<float4*>(st + so0 + lo_off) = l;
            }
            if (ok1) {
              split4(v[i][1], h, l);
              *reinterpret_cast<float4*>(st + so1) = h;
              *reinterpret_cast<float4*>(st + so1 + lo_off) = l;
            }
            if (c + PFR < nch) {
              const int ce = (c + PFR + rot < nch) ? c + PFR + rot : c + PFR + rot - nch;
              if (ok0) v[i][0] = ld_cg_f4(g0 + ce * KC);
              if (ok1) v[i][1] = ld_cg_f4(g1 + ce * KC);
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[slot]);
            if (++slot == (uint32_t)nslot) {
              slot = 0;
              ++use;
            }
          }
        }
      }
    }
    if (warp < 4) {
      // ---------------------------- accumulator -> shared memory ------------------------------
      mbar_wait(accum, ps.acc & 1);
      stamp(53);
      tc::tc_fence_after();
      const int row = warp * 32 + lane;  // TMEM lane = stacked activation row
      const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
      float* rr = red + (size_t)row * p.ldred;
      for (int c = 0; c < nvalid; c += 16) {
        float a[16], b[16];
        tc::tmem_ld16(taddr + (uint32_t)c, a);
        tc::tmem_ld16(taddr + (uint32_t)(nh + c), b);
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (c + q < nvalid) rr[c + q] = a[q] + b[q];
      }
      tc::tc_fence_before();
    }
  } else if (warp == 8) {
    // ------------------------------ MMA issuer ------------------------------------------------
    const uint32_t ring_s = smem_u32(ring);
    for (int c = 0; c < nch; ++c) {
      mbar_wait(&ready[slot], use & 1);
      tc::tc_fence_after();
      {
        // the whole warp runs the burst on warp-uniform operands; elect.sync predicates the instructions (lstmp_tc.cuh)
        const uint32_t a0 = ring_s + slot * SLOT_BYTES;
        const int ce = (c + rot < nch) ? c + rot : c + rot - nch;
        const uint32_t b0 = b_addr + (uint32_t)ce * chunk_b;
#pragma unroll
        for (int j = 0; j < KC / 8; ++j) {
          // MMA j covers the 16-byte K chunks 2j and 2j+1 of the 128-byte rows (K = 8 tf32 per instruction)
          const uint64_t da = tc::make_desc_sw128(a0 + 32 * j);
          const uint64_t db = tc::make_desc_sw128(b0 + 32 * j);
          if (tc::elect_one()) tc::mma_tf32(tmem_d, da, db, idesc, (c | j) ? 1u : 0u);
        }
        if (tc::elect_one()) {
          tc::umma_commit(&empty[slot]);             // frees the slot once these MMAs have read it
          if (c == nch - 1) tc::umma_commit(accum);  // accumulator complete
        }
      }
      __syncwarp();
      if (++slot == (uint32_t)nslot) {
        slot = 0;
        ++use;
      }
    }
    tc::tc_fence_before();
  } else if (warp == 9) {
    // ------------------------------ proxy-fence warp ------------------------------------------
    // fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC.  In the loader threads the MEMBAR would wait
    // for their in-flight cp.async copies, in the issuer thread for its in-flight MMAs; this warp has neither, so it
    // relays "slot written" (full) to "slot visible to the async proxy" (ready).
    for (int c = 0; c < nch; ++c) {
      mbar_wait(&full[slot], use & 1);
      tc::fence_async_smem();
      if (lane == 0) tc::mbar_arrive(&ready[slot]);
      if (++slot == (uint32_t)nslot) {
        slot = 0;
        ++use;
      }
    }
  }
  __syncthreads();
  tc::tc_fence_after();
  ps.cc += (uint32_t)nch;
  ps.acc += 1;
}
}  // namespace tcf

template <int LOADER>
__global__ void __launch_bounds__(kThreads, 1) lstmp_fwd_tc_kernel(const __grid_constant__ FwdTcParams p) {
  using namespace tcf;
  extern __shared__ __align__(16) uint8_t smem_raw_tc[];
#ifdef LSTMP_TC_SHARED_SPACE
  // see lstmp_gemm_tc.cu: keeps the shared address space visible (STS/LDS instead of generic ST.E/LD.E); opt-in build
  uint8_t* base = smem_raw_tc + ((1024u - (smem_u32(smem_raw_tc) & 1023u)) & 1023u);
#else
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~uintptr_t(1023));
#endif
  uint8_t* bg = base + p.off_bg;      // gate weight slice:   R/32 tiles of [roundup8(8*cpc) rows][128 B]  (hi rows, then lo)
  uint8_t* bp = base + p.off_bp;      // projection slice:    C/32 tiles of [roundup8(2*rpc) rows][128 B]
  uint8_t* ring = base + p.off_ring;  // nslot x [128 rows][128 B]
  float* red = reinterpret_cast<float*>(base + p.off_red);      // [128][ldred]; aliases the ring (idle in the epilogue)
  uint8_t* stage = base + p.off_stage;                          // PF x [S*8 units x 16 B] raw cp.async landing slots
  float* cprev = reinterpret_cast<float*>(base + p.off_cprev);  // [S*nc]
  float* peep = reinterpret_cast<float*>(base + p.off_peep);    // [3][cpc]
  uint64_t* full = reinterpret_cast<uint64_t*>(base + p.off_bars);
  uint64_t* ready = full + MAX_SLOTS;
  uint64_t* empty = ready + MAX_SLOTS;
  uint64_t* accum = empty + MAX_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int j = blockIdx.x;
  const int C = p.C, R = p.R, S = p.S, T = p.T;
  const int cpc = p.cpc, rpc = p.rpc;
  const int c0 = j * cpc;
  const int nc = max(0, min(cpc, C - c0));  // my cells
  const int r0 = j * rpc;
  const int nr = max(0, min(rpc, R - r0));  // my projection outputs

  if (tid == 0) {
    for (int s = 0; s < p.nslot; ++s) {
      mbar_init(&full[s], LOADER == 2 ? 1 : LOADERS / 32);
      mbar_init(&ready[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }

  // ---- stationary weight slices: split into hi / lo once per launch (row-fastest mapping: conflict-free stores) ----
  if (nc > 0) {
    const int rows = 4 * nc, nk4 = R >> 2;  // hi row = gate*nc + cl (gate order g,i,f,o: LPS.h:234-243)
    for (int u = tid; u < rows * nk4; u += kThreads) {
      const int row = u % rows, k4 = u / rows;
      const int gate = row / nc, cl = row - gate * nc;
      const float4 x = __ldg(reinterpret_cast<const float4*>(p.w_gifo_r + (size_t)(gate * C + c0 + cl) * R) + k4);
      float4 h, l;
      split4(x, h, l);
      uint8_t* t = bg + (size_t)(k4 >> 3) * p.chunk_g;  // tile of K chunk k4/8: [8*cpc rows][128 B], swizzled
      *reinterpret_cast<float4*>(t + tc::sw128_off(row, k4 & 7)) = h;
      *reinterpret_cast<float4*>(t + tc::sw128_off(4 * cpc + row, k4 & 7)) = l;
    }
  }
  if (nr > 0) {
    const int nk4 = C >> 2;
    for (int u = tid; u < nr * nk4; u += kThreads) {
      const int row = u % nr, k4 = u / nr;
      const float4 x = __ldg(reinterpret_cast<const float4*>(p.w_r_m + (size_t)(r0 + row) * C) + k4);
      float4 h, l;
      split4(x, h, l);
      uint8_t* t = bp + (size_t)(k4 >> 3) * p.chunk_p;
      *reinterpret_cast<float4*>(t + tc::sw128_off(row, k4 & 7)) = h;
      *reinterpret_cast<float4*>(t + tc::sw128_off(rpc + row, k4 & 7)) = l;
    }
  }
  tc::fence_async_smem();

  // ---- history: c_0 of my cells -> smem and cbuf block 0; r_0 of my columns -> rbuf block 0 (LPS.h:231) ----
  for (int idx = tid; idx < S * nc; idx += kThreads) {
    int s = idx / nc, cl = idx - s * nc;
    float v = p.state_c[(size_t)s * C + c0 + cl];
    cprev[idx] = v;
    p.cbuf[(size_t)s * C + c0 + cl] = v;
  }
  for (int idx = tid; idx < S * nr; idx += kThreads) {
    int s = idx / nr, n = idx - s * nr;
    p.rbuf[(size_t)s * R + r0 + n] = p.state_r[(size_t)s * R + r0 + n];
  }
  for (int cl = tid; cl < nc; cl += kThreads) {
    peep[cl] = p.p_i[c0 + cl];
    peep[cpc + cl] = p.p_f[c0 + cl];
    peep[2 * cpc + cl] = p.p_o[c0 + cl];
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t idesc_g = tc::idesc_tf32(128, p.n_g), idesc_p = tc::idesc_tf32(128, p.n_p);
  const uint32_t bg_s = smem_u32(bg), bp_s = smem_u32(bp);
  const int ldred = (int)p.ldred;

  GroupBarrier gb;
  gb.init(p.bar, p.bar_base, (unsigned)p.nctas, (p.dbg & 1) != 0);
  stamp_begin((p.dbg & 4) && blockIdx.x == 0);
  __syncthreads();
  stamp(1);
  Pipe ps{0u, 0u};

  for (int tt = 0; tt < T; ++tt) {
    stamp(10);
    // ================= phase 1: gates + cell update for my cells ==========================
    if (nc > 0) {
      // this thread's first element's x-part pre-activations (input GEMM + bias): in flight during the product
      float xg = 0.f, xi = 0.f, xf = 0.f, xo = 0.f;
      if (tid < S * nc) {
        int s = tid / nc, cl = tid - s * nc;
        const float* gp = p.gifo + (size_t)(tt * S + s) * (4 * C) + c0 + cl;
        xg = gp[0];
        xi = gp[C];
        xf = gp[2 * C];
        xo = gp[3 * C];
      }
      // r_{t-1}: carried state for the first frame of the chunk, else rbuf block tt
      const float* X = (tt == 0) ? p.state_r : p.rbuf + (size_t)tt * S * R;
      // gifo(t) += r(t-1) * W_gifo_r^T                                       (LPS.h:275)
      tc_product<LOADER>(p, ps, X, R, R, bg_s, p.chunk_g, idesc_g, tmem_base + COL_G, 4 * cpc, 4 * nc, ring, full, ready, empty, accum,
                 red, stage);
      stamp(11);
      for (int idx = tid; idx < S * nc; idx += kThreads) {
        int s = idx / nc, cl = idx - s * nc;
        size_t row = (size_t)tt * S + s;
        float* gp = p.gifo + row * (4 * C) + c0 + cl;
        if (idx >= kThreads) {
          xg = gp[0];
          xi = gp[C];
          xf = gp[2 * C];
          xo = gp[3 * C];
        }
        const float* rh = red + s * ldred + cl;        // hi-activation rows
        const float* rl = red + (S + s) * ldred + cl;  // lo-activation rows
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
        gp[0] = gg;
        gp[C] = gi;
        gp[2 * C] = gf;
        gp[3 * C] = go;
        p.cbuf[(row + S) * C + c0 + cl] = c;
        p.hbuf[row * C + c0 + cl] = h;
        p.mbuf[row * C + c0 + cl] = m;
        cprev[idx] = c;
      }
    }
    stamp(20);
    gb.sync();
    stamp(21);
    // ================= phase 2: projection r(t) = m(t) * W_r_m^T for my columns (LPS.h:312) ==
    if (nr > 0) {
      tc_product<LOADER>(p, ps, p.mbuf + (size_t)tt * S * C, C, C, bp_s, p.chunk_p, idesc_p, tmem_base + COL_P, rpc, nr, ring,
                 full, ready, empty, accum, red, stage);
      stamp(22);
      for (int idx = tid; idx < S * nr; idx += kThreads) {
        int s = idx / nr, n = idx - s * nr;
        float v = red[s * ldred + n] + red[(S + s) * ldred + n];
        size_t row = (size_t)tt * S + s;
        p.rbuf[(row + S) * R + r0 + n] = v;
        p.out[row * p.ld_out + r0 + n] = v;                         // :328
        if (tt == T - 1) p.state_r[(size_t)s * R + r0 + n] = v;     // :331
      }
    }
    stamp(30);
    if (tt + 1 < T) gb.sync();
    stamp(31);
  }
  // prev_nnet_state_ <- last frame (LPS.h:331): c part
  for (int idx = tid; idx < S * nc; idx += kThreads) {
    int s = idx / nc, cl = idx - s * nc;
    p.state_c[(size_t)s * C + c0 + cl] = cprev[idx];
  }
  stamp_flush(p.dbg_stamps);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool fwd_tc_plan(int C, int R, int S, int nctas, size_t smem_limit, int loader, FwdTcParams* p, size_t* smem_bytes) {
  using namespace tcf;
  if (S < 8 || S > 64 || (S & 7)) return false;  // A tile: 2*S stacked rows <= 128 = the MMA's M; lo rows at row S
  if (C % KC || R % KC || nctas < 1) return false;
  const int cpc = (C + nctas - 1) / nctas, rpc = (R + nctas - 1) / nctas;
  const int n_g = (8 * cpc + 15) & ~15, n_p = (2 * rpc + 15) & ~15;
  if (n_g > 256 || n_p > 256) return false;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 1023) & ~size_t(1023);
    return (unsigned)o;
  };
  p->cpc = cpc;
  p->rpc = rpc;
  p->n_g = n_g;
  p->n_p = n_p;
  p->nctas = nctas;
  p->chunk_g = (unsigned)(((8 * cpc + 7) & ~7) * 128);  // one K chunk of the stacked gate slice (multiple of 1024 B)
  p->chunk_p = (unsigned)(((2 * rpc + 7) & ~7) * 128);
  p->off_bg = take((size_t)(R / KC) * p->chunk_g);
  p->off_bp = take((size_t)(C / KC) * p->chunk_p);
  // the MMA reads N (>= actual) weight rows per slab: the over-read of the last slabs lands in the ring that
  // follows (any finite-or-not garbage only reaches accumulator columns nobody reads)
  p->off_ring = (unsigned)off;
  const int ldred = ((4 * cpc > rpc ? 4 * cpc : rpc) | 1);
  p->ldred = (unsigned)ldred;
  if (loader < 0 || loader > 2) return false;
  p->loader = loader;
  const size_t stage_total = loader == 1 ? (size_t)PF * S * SLABS * 16 : 0;
  const size_t tail = ((stage_total + 1023) & ~size_t(1023)) + (((size_t)S * cpc * 4 + 1023) & ~size_t(1023)) +
                      1024 /* peepholes */ + 1024 /* barriers + tmem slot */;
  const size_t reserve = 1024 /* base alignment */ + (size_t)static_smem_reserve();
  if (smem_limit < off + tail + reserve + (size_t)2 * SLOT_BYTES) return false;
  int nslot = (int)((smem_limit - off - tail - reserve) / SLOT_BYTES);
  if (nslot > MAX_SLOTS) nslot = MAX_SLOTS;
  p->nslot = nslot;
  p->slot_bytes = SLOT_BYTES;
  off += (size_t)nslot * SLOT_BYTES;
  if ((size_t)128 * ldred * 4 > (size_t)nslot * SLOT_BYTES) return false;
  p->off_red = p->off_ring;  // the accumulator exchange buffer aliases the ring, which is idle during the epilogue
  p->off_stage = take(stage_total);
  p->off_cprev = take((size_t)S * cpc * 4);
  p->off_peep = take((size_t)3 * cpc * 4);
  p->off_bars = take(256);
  *smem_bytes = off + 1024;
  return *smem_bytes + (size_t)static_smem_reserve() <= smem_limit;
}

cudaError_t fwd_tc_set_smem_limit(size_t bytes) {
  // The attribute is per device and per function and several engines of different shapes may share the process:
  // only ever raise it (as set_kernel_smem_limits does for the FFMA kernels).
  static size_t cur[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  dev &= 63;
  if (bytes <= cur[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute((const void*)lstmp_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((const void*)lstmp_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((const void*)lstmp_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)bytes);
  if (e != cudaSuccess) return e;
  cur[dev] = bytes;
  return cudaSuccess;
}

cudaError_t launch_fwd_tc(const FwdTcParams& p, size_t smem_bytes, cudaStream_t stream) {
  void* args[] = {(void*)&p};
  dim3 grid(p.nctas), block(kThreads);
  const void* fn = p.loader == 1 ? (const void*)lstmp_fwd_tc_kernel<1>
                   : p.loader == 2 ? (const void*)lstmp_fwd_tc_kernel<2> : (const void*)lstmp_fwd_tc_kernel<0>;
  return cudaLaunchCooperativeKernel(fn, grid, block, args, smem_bytes, stream);
}

}  // namespace lstmp
