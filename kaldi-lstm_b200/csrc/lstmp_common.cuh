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
// Monotonic counter: every CTA adds 1 per barrier; barrier k (1-based) of this launch is
// complete when counter - base >= k * nctas.  Release/acquire through __threadfence() as
// cooperative_groups::grid_group::sync() does; wrap-safe signed comparison.
// ---------------------------------------------------------------------------------------
struct GroupBarrier {
  unsigned* counter;
  unsigned target;  // thread 0's running target
  unsigned nctas;
  __device__ __forceinline__ void init(unsigned* c, unsigned base, unsigned n) {
    counter = c;
    target = base;
    nctas = n;
  }
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += nctas;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(counter) : "memory");
      } while (static_cast<int>(v - target) < 0);
      __threadfence();
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

// ---------------------------------------------------------------------------------------
// Skinny product with a stationary weight slice (the recurrent / projection steps):
//
//   red[s*ldred + n] = sum_{k<K} X[s][k] * W[n][k]      s < Sg, n < Nc
//
// X is [Sg x K] in global memory (written by other CTAs before the preceding group barrier)
// and is streamed through a double-buffered shared-memory ring with cp.async in K-chunks;
// W is [Nc x K] resident in shared memory (row stride ldw).  Each thread owns a 4x4
// (streams x columns) register tile; the K range of a chunk is split over `ksplit` adjacent
// lanes whose partial sums are combined with warp shuffles (ksplit is a power of two <= 32).
// All kThreads threads must call this (it contains __syncthreads and full-warp shuffles).
// ---------------------------------------------------------------------------------------
struct SkinnyMap {
  int n_s_tiles, n_n_tiles, tiles, ksplit;
  __device__ __forceinline__ void make(int Sg, int Nc) {
    n_s_tiles = ceil_div(Sg, 4);
    n_n_tiles = ceil_div(Nc, 4);
    tiles = n_s_tiles * n_n_tiles;
    int ks = kThreads / tiles;
    ksplit = ks >= 1 ? pow2_floor(ks < 32 ? ks : 32) : 1;
  }
};

__device__ __forceinline__ int xbuf_ld(int KC, int ksplit) { return KC + (ksplit < 8 ? 4 * ksplit : 4); }

__device__ __forceinline__ void skinny_load_chunk(float* xs, int ldx, const float* __restrict__ Xg, size_t ldX,
                                                  int Sg, int k0, int L) {
  const int q_per_row = L >> 2;
  const int total = Sg * q_per_row;
  for (int idx = threadIdx.x; idx < total; idx += kThreads) {
    int s = idx / q_per_row, q = idx - s * q_per_row;
    cp_async16(xs + s * ldx + 4 * q, Xg + (size_t)s * ldX + k0 + 4 * q);
  }
}

static __device__ __noinline__ void skinny_gemm(const float* __restrict__ Xg, size_t ldX, int K, int Sg,
                                         const float* __restrict__ Ws, int ldw, int Nc, float* xbuf, int KC,
                                         float* red, int ldred) {
  SkinnyMap m;
  m.make(Sg, Nc);
  const int ldx = xbuf_ld(KC, m.ksplit);
  const int tid = threadIdx.x;
  const int pass_tiles = kThreads / m.ksplit;  // tiles processed concurrently
  const int npass = ceil_div(m.tiles, pass_tiles);
  const int kq = tid % m.ksplit;
  const int nchunks = ceil_div(K, KC);

  for (int pass = 0; pass < npass; ++pass) {
    const int tile = pass * pass_tiles + tid / m.ksplit;
    const bool active = tile < m.tiles;
    const int s_tile = active ? tile % m.n_s_tiles : 0;
    const int n_tile = active ? tile / m.n_s_tiles : 0;
    int srow[4], ncol[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int s = s_tile + i * m.n_s_tiles;
      srow[i] = (s < Sg ? s : Sg - 1) * ldx;
      int n = n_tile + i * m.n_n_tiles;
      ncol[i] = (n < Nc ? n : Nc - 1) * ldw;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    skinny_load_chunk(xbuf, ldx, Xg, ldX, Sg, 0, (K < KC ? K : KC));
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
      const int k0 = c * KC;
      const int L = (K - k0 < KC ? K - k0 : KC);
      if (c + 1 < nchunks) {
        const int k1 = k0 + KC;
        skinny_load_chunk(xbuf + ((c + 1) & 1) * Sg * ldx, ldx, Xg, ldX, Sg, k1, (K - k1 < KC ? K - k1 : KC));
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      if (active) {
        const float* xs = xbuf + (c & 1) * Sg * ldx;
        const float* ws = Ws + k0;
        const int nq = L >> 2;
#pragma unroll 2
        for (int q = kq; q < nq; q += m.ksplit) {
          float4 xv[4], wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + srow[i] + 4 * q);
#pragma unroll
          for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4*>(ws + ncol[j] + 4 * q);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[i][j] = fmaf(xv[i].x, wv[j].x, acc[i][j]);
              acc[i][j] = fmaf(xv[i].y, wv[j].y, acc[i][j]);
              acc[i][j] = fmaf(xv[i].z, wv[j].z, acc[i][j]);
              acc[i][j] = fmaf(xv[i].w, wv[j].w, acc[i][j]);
            }
        }
      }
      __syncthreads();
    }
    // combine the ksplit partial sums held by adjacent lanes
    for (int off = m.ksplit >> 1; off >= 1; off >>= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], off);
    }
    if (active && kq == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int s = s_tile + i * m.n_s_tiles;
        if (s < Sg) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int n = n_tile + j * m.n_n_tiles;
            if (n < Nc) red[s * ldred + n] = acc[i][j];
          }
        }
      }
    }
  }
  __syncthreads();
}

}  // namespace lstmp
