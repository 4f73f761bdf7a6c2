// Common device helpers for the B200-native LstmProjectedStreams engine (sm_100a only).
//
// Nothing here is a port: the reference (google/nnet/bd-nnet-lstm-projected-streams.h) only
// calls CuMatrix methods; these helpers serve the fused persistent per-chunk kernels that
// replace that call sequence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lstmp {

constexpr int kThreads = 384;      // 12 warps per persistent CTA, 1 CTA per SM
constexpr int kMaxGroups = 8;      // independent stream groups (each with its own grid barrier)
constexpr float kCellClip = 50.0f; // LPS.h:296-297

// ---------------------------------------------------------------------------------------
// Small PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared, L2 only (.cg): used for every operand that another CTA
// wrote earlier in the same launch, so no stale L1 line can be observed.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_cg_f(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
  return v;
}

// --- mbarrier + TMA bulk copy (1-D cp.async.bulk, SASS UBLKCP): weight slices are staged into
// shared memory once per launch and reused for every timestep of the chunk.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------------
// Group-wide barrier between the co-resident CTAs of one stream group (cooperative launch).
// One monotonic arrival counter per group: every CTA does one release-ordered RED (+1) and one
// thread polls the counter with RELAXED loads (an acquire load costs an L1 invalidate per poll),
// followed by a single acquire fence.  Barrier k of the launch completes when
// counter - base >= k * nctas (wrap-safe).  Measured ~1.5 us on 148 CTAs; a per-CTA flag array
// polled by every CTA was slower (hot-line contention), see DESIGN.md.
// ---------------------------------------------------------------------------------------
struct GroupBarrier {
  unsigned* counter;
  unsigned target;  // thread 0's running target
  unsigned nctas;
  bool off;
  __device__ __forceinline__ void init(unsigned* c, unsigned base, unsigned n, bool disabled) {
    counter = c;
    target = base;
    nctas = n;
    off = disabled;
  }
  __device__ __forceinline__ void sync() {
    __syncthreads();  // all of this CTA's global writes are ordered before thread 0's release
    if (threadIdx.x == 0) {
      target += nctas;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
      if (!off) {
        unsigned v;
        do {
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(counter) : "memory");
        } while (static_cast<int>(v - target) < 0);
        asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
      }
    }
    __syncthreads();
  }
};

// ---------------------------------------------------------------------------------------
// Activations.  The reference's CPU path uses Kaldi's overflow-safe forms (
// VectorBase::Sigmoid/Tanh); __expf is accurate to ~2 ulp over the ranges that matter and the
// saturating branches below reproduce the same limits, well inside the 1e-4 tolerance.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_fast(float x) {
  // 1/(1+e^-x); for x << 0, e^-x -> inf and the quotient -> 0 as required.
  return __fdividef(1.0f, 1.0f + __expf(-x));
}
__device__ __forceinline__ float tanhf_fast(float x) {
  // tanh(x) = 1 - 2/(1+e^{2x}); clamp the exponent so e^{2x} stays finite.
  float e = __expf(2.0f * fminf(fmaxf(x, -20.0f), 20.0f));
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int pow2_floor(int x) {
  int p = 1;
  while (p * 2 <= x) p *= 2;
  return p;
}

// Debug time stamps (build with -DLSTMP_STAMPS and run with LSTMP_B200_DEBUG & 4): thread 0 of CTA 0
// appends (tag, clock64) pairs to a shared-memory log that is flushed to global memory at kernel end.
#ifdef LSTMP_STAMPS
constexpr int kMaxStamps = 384;
__shared__ long long s_stamp_log[2 * kMaxStamps];
__shared__ int s_stamp_n;
__shared__ int s_stamp_on;
__device__ __forceinline__ void stamp(int tag) {
  // thread 0 everywhere; in the TMA-fed loops also the sync / producer thread (256) and the MMA issuer's lane 0 (288)
  if ((threadIdx.x == 0 || ((threadIdx.x == 256 || threadIdx.x == 288) && tag >= 200)) && s_stamp_on) {
    int n = atomicAdd(&s_stamp_n, 1);
    if ((unsigned)n < (unsigned)kMaxStamps) {
      s_stamp_log[2 * n] = tag;
      s_stamp_log[2 * n + 1] = clock64();
    }
  }
}
__device__ __forceinline__ void stamp_begin(bool on) {
  if (threadIdx.x == 0) {
    s_stamp_on = on ? 1 : 0;
    s_stamp_n = 0;
  }
}
__device__ __forceinline__ void stamp_flush(long long* dst) {
  if (threadIdx.x == 0 && s_stamp_on && dst) {
    long long base = dst[0];
    int n = s_stamp_n < kMaxStamps ? s_stamp_n : kMaxStamps;
    for (int i = 0; i < n && base + i < 1024; ++i) {
      dst[2 + 2 * (base + i)] = s_stamp_log[2 * i];
      dst[3 + 2 * (base + i)] = s_stamp_log[2 * i + 1];
    }
    dst[0] = base + n < 1024 ? base + n : 1024;
  }
}
constexpr int kStaticSmemReserve = 2 * kMaxStamps * 8 + 128;
#else
__device__ __forceinline__ void stamp(int) {}
__device__ __forceinline__ void stamp_begin(bool) {}
__device__ __forceinline__ void stamp_flush(long long*) {}
constexpr int kStaticSmemReserve = 128;
#endif

// ---------------------------------------------------------------------------------------
// Skinny product with a stationary weight slice (the recurrent / projection steps):
//
//   red[s*ldred + n] = sum_{k<K} X[s][k] * W[n][k]      s < Sg, n < Nc
//
// X is [Sg x K] in global memory (written by other CTAs before the preceding group barrier) and is
// streamed through a ring of shared-memory slots of [Sg x kch] floats with cp.async; ALL chunks that fit
// are put in flight at once so the L2 round trip is paid once per phase.  W is [Nc x K] resident in
// shared memory (row stride ldw).  Each thread owns a 4x4 (streams x columns) register tile; the K
// range of a chunk is split over `ksplit` adjacent lanes whose partial sums are combined with warp
// shuffles (ksplit is a power of two <= 32).  The thread->tile mapping (a "plan") depends only on
// (Sg, Nc, K) and is computed once per launch, outside the time loop.
// All kThreads threads must call skinny_gemm (it contains __syncthreads and full-warp shuffles).
// ---------------------------------------------------------------------------------------
constexpr int kXbufPadMax = 16;
__device__ __forceinline__ int xbuf_ld(int kch, int ksplit) { return kch + (ksplit < 8 ? 4 * ksplit : 4); }

struct SkinnyPlan {          // lives in shared memory; uniform fields + one packed word per thread
  int Sg, Nc, K;
  int n_s_tiles, n_n_tiles, ksplit, npass;
  int tsz;                   // streams per thread tile: 8 (x 4 columns) for groups of >= 16 streams, else 4
  int ts, tn, nbs, nblocks;  // a warp covers a ts x tn block of tiles; nbs blocks along s; nblocks total
  int kch, nchunks, nslots, ldx, slot_floats;  // ring: slot_floats = Sg*ldx; panel: chunk c at column c*kch
  int panel;                                   // 1: the whole [Sg x K] panel is resident (no slot reuse)
  unsigned thr[kThreads];    // bit 31 active | kq [20,26) | n_tile [10,20) | s_tile [0,10)   (pass 0)
};

// Tile coordinates of (warp-block wb, lane).  A warp holds 32/ksplit tiles arranged ts x tn so that the
// distinct shared-memory rows touched by one LDS.128 stay few for BOTH operands (fewer wavefronts).
__device__ __forceinline__ void skinny_tile_of(const SkinnyPlan* pl, int wb, int lane, bool* active, int* s_tile,
                                               int* n_tile) {
  const int t = lane / pl->ksplit;
  const int ds = t % pl->ts, dn = t / pl->ts;
  const int bs = wb % pl->nbs, bn = wb / pl->nbs;
  *s_tile = bs * pl->ts + ds;
  *n_tile = bn * pl->tn + dn;
  *active = wb < pl->nblocks && *s_tile < pl->n_s_tiles && *n_tile < pl->n_n_tiles && pl->Nc > 0;
  if (!*active) *s_tile = *n_tile = 0;
}

// cap_floats: capacity of the shared all-gather buffer.  If the whole [Sg x K] panel fits, every chunk gets its
// own columns and all of them are put in flight at once; otherwise a ring of [Sg x kch] slots is used.
__device__ __forceinline__ void skinny_make_plan(SkinnyPlan* pl, int Sg, int Nc, int K, int cap_floats) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    // The loop is bound by the SM-wide LDS.128 issue rate (one non-uniform 512-byte request per ~4 cycles).  8x4
    // tiles (12 LDS.128 per 128 FMAs instead of 8 per 64) cut the product time by 15 % but doubled the shuffle
    // reduction and lost overall (fwd 442 us vs 412 us at S=64), so 4x4 stays the default; see DESIGN.md.
    const int tsz = 4;
    pl->tsz = tsz;
    const int n_s_tiles = ceil_div(Sg, tsz), n_n_tiles = ceil_div(Nc > 0 ? Nc : 1, 4);
    const int tiles = n_s_tiles * n_n_tiles;
    int ks = kThreads / tiles;
    const int ksplit = ks >= 1 ? pow2_floor(ks < 32 ? ks : 32) : 1;
    const int tpw = 32 / ksplit;  // tiles per warp (power of two)
    // ts x tn = tpw, as square as the tile grid allows
    int ts = 1;
    while (ts * ts < tpw) ts *= 2;             // ts = ceil-pow2(sqrt(tpw))
    if (ts > tpw) ts = tpw;
    while (ts > 1 && ts > pow2_floor(n_s_tiles) * 2) ts /= 2;
    int tn = tpw / ts;
    while (tn > 1 && tn > pow2_floor(n_n_tiles) * 2 && ts < tpw) { tn /= 2; ts *= 2; }
    const int nbs = ceil_div(n_s_tiles, ts), nbn = ceil_div(n_n_tiles, tn);
    pl->Sg = Sg; pl->Nc = Nc; pl->K = K;
    pl->n_s_tiles = n_s_tiles; pl->n_n_tiles = n_n_tiles; pl->ksplit = ksplit;
    pl->ts = ts; pl->tn = tn; pl->nbs = nbs; pl->nblocks = nbs * nbn;
    pl->npass = ceil_div(nbs * nbn, kThreads / 32);
    const int pad = (ksplit < 8 ? 4 * ksplit : 4);
    const int kround = (K + 31) & ~31;
    if ((long long)Sg * (kround + pad) <= cap_floats && Sg <= 16) {
      int kch = kround;  // small stream groups: the whole K in one chunk, no per-chunk loop
      while (ceil_div(K, kch) > 8) kch += 128;  // at most 8 cp.async groups in flight
      pl->panel = 1; pl->kch = kch; pl->nchunks = ceil_div(K, kch); pl->nslots = pl->nchunks;
      pl->ldx = kround + pad; pl->slot_floats = 0;
    } else {
      int kch = 256;
      if ((long long)2 * Sg * (kch + pad) > cap_floats) kch = 128;
      int ns = cap_floats / (Sg * (kch + pad));
      pl->panel = 0; pl->kch = kch; pl->nchunks = ceil_div(K, kch); pl->nslots = ns > 8 ? 8 : (ns < 1 ? 1 : ns);
      pl->ldx = kch + pad; pl->slot_floats = Sg * (kch + pad);
    }
  }
  __syncthreads();
  bool active;
  int s_tile, n_tile;
  skinny_tile_of(pl, tid >> 5, tid & 31, &active, &s_tile, &n_tile);
  pl->thr[tid] = ((active ? 1u : 0u) << 31) | ((unsigned)((tid & 31) % pl->ksplit) << 20) | ((unsigned)n_tile << 10) |
                 (unsigned)s_tile;
}

__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    default: cp_async_wait<7>(); break;
  }
}

// Packed 2-wide FP32 FMA (Blackwell FFMA2, PTX fma.rn.f32x2): two fp32 FMAs per issued instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

template <int TS>
static __device__ __noinline__ void skinny_gemm_t(const SkinnyPlan* __restrict__ pl, const float* __restrict__ Xg,
                                                  size_t ldX, const float* __restrict__ Ws, int ldw, float* xbuf,
                                                  float* red, int ldred) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Sg = pl->Sg, Nc = pl->Nc, K = pl->K;
  if (Nc <= 0) return;  // CTA-uniform
  const int ksplit = pl->ksplit, ldx = pl->ldx, slot_floats = pl->slot_floats;
  const int kch = pl->kch, nchunks = pl->nchunks, nslots = pl->nslots, npass = pl->npass;
  const int n_s_tiles = pl->n_s_tiles, n_n_tiles = pl->n_n_tiles;
  const bool panel = pl->panel != 0;
  const unsigned packed = pl->thr[tid];
  const int kq = (packed >> 20) & 63;

  int issue_slot = 0, issue_k0 = 0;
  auto issue = [&]() {  // next chunk -> its panel columns / the next ring slot
    const int L = (K - issue_k0 < kch ? K - issue_k0 : kch) >> 2;  // 16-byte units per row
    float* dst = panel ? xbuf + issue_k0 : xbuf + issue_slot * slot_floats;
    const float* src = Xg + issue_k0;
    for (int s = warp; s < Sg; s += kThreads / 32) {
      float* d = dst + s * ldx;
      const float* g = src + (size_t)s * ldX;
      for (int u = lane; u < L; u += 32) cp_async16(d + 4 * u, g + 4 * u);
    }
    cp_async_commit();
    issue_k0 += kch;
    if (++issue_slot == nslots) issue_slot = 0;
  };

  for (int pass = 0; pass < npass; ++pass) {
    bool active;
    int s_tile, n_tile;
    if (pass == 0) {
      active = (packed >> 31) != 0;
      s_tile = packed & 1023;
      n_tile = (packed >> 10) & 1023;
    } else {
      skinny_tile_of(pl, pass * (kThreads / 32) + warp, lane, &active, &s_tile, &n_tile);
    }
    issue_slot = 0;
    issue_k0 = 0;
    int issued = 0;
    stamp(100);
    for (; issued < nslots && issued < nchunks; ++issued) issue();
    stamp(101);
    // (address set-up below overlaps with the L2 round trip of the chunks just issued)
    // byte addresses (shared window) of this thread's 4 X rows and 4 W rows, at its first quad
    uint32_t xa[TS], wa[4];
#pragma unroll
    for (int i = 0; i < TS; ++i) {
      int s = s_tile + i * n_s_tiles;
      xa[i] = smem_u32(xbuf) + (uint32_t)(((s < Sg ? s : Sg - 1) * ldx + 4 * kq) * 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n = n_tile + i * n_n_tiles;
      wa[i] = smem_u32(Ws) + (uint32_t)(((n < Nc ? n : Nc - 1) * ldw + 4 * kq) * 4);
    }
    const uint32_t qstep = 16u * ksplit;
    // acc2[i][j] = (sum over even k, sum over odd k) of X[s_i][k] * W[n_j][k]: the natural (x,y) / (z,w) register
    // pairs of the 128-bit loads feed FFMA2 directly, no operand duplication.
    f32x2 acc2[TS][4];
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;

    int slot = 0;
    for (int c = 0; c < nchunks; ++c) {
      cp_async_wait_dyn(issued - c - 1);
      __syncthreads();  // chunk c landed for everyone; everyone is done with chunk c-1
      stamp(110 + c);
      if (c >= 1 && issued < nchunks) {
        issue();  // ring only: into the slot chunk c-1 just vacated
        ++issued;
      }
      if (active) {
        const int k0 = c * kch;
        const int nq = (K - k0 < kch ? K - k0 : kch) >> 2;
        const uint32_t xoff = (uint32_t)((panel ? k0 : slot * slot_floats) * 4), woff = (uint32_t)(k0 * 4);
        // (a hand software-pipelined variant of this loop measured slower than letting ptxas schedule it)
        uint32_t xp[TS], wp[4];
#pragma unroll
        for (int i = 0; i < TS; ++i) xp[i] = xa[i] + xoff;
#pragma unroll
        for (int j = 0; j < 4; ++j) wp[j] = wa[j] + woff;
#pragma unroll 2
        for (int q = kq; q < nq; q += ksplit) {
          f32x2 xl[TS], xh[TS], wl[4], wh[4];
#pragma unroll
          for (int i = 0; i < TS; ++i) {
            const float4 v = lds128(xp[i]);
            xp[i] += qstep;
            xl[i] = pack2(v.x, v.y);
            xh[i] = pack2(v.z, v.w);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = lds128(wp[j]);
            wp[j] += qstep;
            wl[j] = pack2(v.x, v.y);
            wh[j] = pack2(v.z, v.w);
          }
#pragma unroll
          for (int i = 0; i < TS; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc2[i][j] = ffma2(xl[i], wl[j], acc2[i][j]);
              acc2[i][j] = ffma2(xh[i], wh[j], acc2[i][j]);
            }
        }
      }
      if (++slot == nslots) slot = 0;
    }
    stamp(130);
    float acc[TS][4];
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float e, o;
        unpack2(acc2[i][j], e, o);
        acc[i][j] = e + o;
      }
    // combine the ksplit partial sums held by adjacent lanes
    for (int off = ksplit >> 1; off >= 1; off >>= 1) {
#pragma unroll
      for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], off);
    }
    if (active && kq == 0) {
#pragma unroll
      for (int i = 0; i < TS; ++i) {
        int s = s_tile + i * n_s_tiles;
        if (s < Sg) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int n = n_tile + j * n_n_tiles;
            if (n < Nc) red[s * ldred + n] = acc[i][j];
          }
        }
      }
    }
    __syncthreads();  // the all-gather buffer and red[] are safe to reuse
    stamp(131);
  }
}

static __device__ __forceinline__ void skinny_gemm(const SkinnyPlan* __restrict__ pl, const float* __restrict__ Xg,
                                                   size_t ldX, const float* __restrict__ Ws, int ldw, float* xbuf,
                                                   float* red, int ldred) {
  if (pl->tsz == 8) skinny_gemm_t<8>(pl, Xg, ldX, Ws, ldw, xbuf, red, ldred);  // CTA-uniform
  else skinny_gemm_t<4>(pl, Xg, ldX, Ws, ldw, xbuf, red, ldred);
}

}  // namespace lstmp
