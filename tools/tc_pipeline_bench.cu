// Microbenchmark of the tcgen05 forward loop's per-chunk pipeline in isolation (the open "600 cycles per
// 8 KB chunk" question).  One CTA per SM, every CTA gathers the SAME [64 x K] fp32 block from L2 (the all-gather
// pattern), 8 loader warps split it into hi/lo and store it into a ring of SWIZZLE_128B [128 x 128 B] slots, a relay
// warp executes the proxy fence, warp 8 issues 4 MMAs (M = 128, N = 48, K = 8) per chunk and commits the slot's
// "empty" barrier -- the structure of tc_product() in kaldi-lstm_b200/csrc/lstmp_recurrent_tc.cu -- with switches:
//
//   GENERIC  : ring / landing-slot pointers derived through uintptr_t (generic LD.E/ST.E, what round 1 shipped) vs
//              pointer arithmetic on the __shared__ array (LDS/STS)
//   staged   : cp.async landing slots + read-back (1) vs LDG.128 register prefetch (0)
//   mma      : issue the MMAs (1) or commit immediately (0: pure loader + handshake cost)
//   nslot    : ring depth
//   relay    : proxy fence in a relay warp (1, what the kernel does), in the issuer warp after its acquire-wait (0),
//              in THREE relay warps taking the chunks round-robin (3), or no fence at all (-1: timing only, the MMAs may
//              then read stale data) -- is one fence.proxy.async per chunk on a single warp the 600-cycle limiter?
//   N        : MMA N (48 = gate product, 16 = projection)
//   wpc      : WARP-PER-CHUNK loaders (1): warp w loads, splits and stores chunks c = w (mod 8) on its own (16 units per
//              lane, full-barrier count 1), so eight chunks are in progress at once -- in the shipped layout (0) every
//              loader warp takes part in every chunk and a chunk cannot finish faster than one warp's dependent chain
//              wait -> read back -> split -> store -> arrive
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I kaldi-lstm_b200/csrc -o tools/_build/tc_pipeline_bench tools/tc_pipeline_bench.cu
#include <cstdio>
#include <vector>
#include "lstmp_tc.cuh"
using namespace lstmp;
using namespace lstmp::tc;

constexpr int KC = 32, S = 64, LOADERS = 256, PF = 3, MAXSLOT = 8;
constexpr uint32_t SLOT = 128 * 128, STAGE = S * 8 * 16;

__device__ __forceinline__ void split4(float4 x, float4& h, float4& l) {
  h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
  h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
  l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
}

template <bool GENERIC>
__global__ void __launch_bounds__(384, 1) pipe(const float* __restrict__ X, int K, int reps, int staged, int mma,
                                               int nslot, int relay, int N, int wpc, long long* out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* base;
  if (GENERIC) base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  else base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wts = base;                       // one 48-row x 128 B tile, reused for every chunk (timing only)
  uint8_t* ring = base + 8 * 1024;
  uint8_t* stage = ring + MAXSLOT * SLOT;    // PF landing slots
  uint64_t* full = reinterpret_cast<uint64_t*>(stage + PF * STAGE);
  uint64_t* ready = full + MAXSLOT;
  uint64_t* empty = ready + MAXSLOT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + MAXSLOT);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (int s = 0; s < nslot; ++s) { mbar_init(&full[s], wpc ? 1 : LOADERS / 32); mbar_init(&ready[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  for (int i = tid; i < 8 * 1024 / 4; i += 384) reinterpret_cast<float*>(wts)[i] = 0.001f;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int nch = K / KC;
  uint32_t cc = 0;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    uint32_t slot = cc % (uint32_t)nslot, use = cc / (uint32_t)nslot;
    if (warp < 8 && wpc) {
      // warp-per-chunk: this warp owns chunks c = warp (mod 8); lane l owns units u = l + 32 i (row u >> 3, 16-byte
      // column u & 7), loads them with LDG.128, splits and stores hi / lo, then lane 0 arrives (count 1)
      const int kc = lane & 7;
      for (int c = 0; c < nch; ++c) {
        if ((c & 7) == warp) {
          float4 x[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = ld_cg_f4(X + (size_t)((lane >> 3) + 4 * i) * K + c * KC + 4 * kc);
          if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
          uint8_t* st = ring + (size_t)slot * SLOT;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = (lane >> 3) + 4 * i;
            float4 h, l;
            split4(x[i], h, l);
            *reinterpret_cast<float4*>(st + sw128_off(r, kc)) = h;
            *reinterpret_cast<float4*>(st + sw128_off(r, kc) + S * 128) = l;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[slot]);
        }
        if (++slot == (uint32_t)nslot) { slot = 0; ++use; }
      }
    } else if (warp < 8) {
      const int r0 = tid >> 3, r1 = r0 + 32, kc = tid & 7;
      const float* g0 = X + (size_t)r0 * K + 4 * kc;
      const float* g1 = X + (size_t)r1 * K + 4 * kc;
      const uint32_t so0 = sw128_off(r0, kc), so1 = sw128_off(r1, kc), lo = S * 128;
      uint8_t* my0 = stage + (size_t)tid * 16;
      uint8_t* my1 = stage + (size_t)(tid + LOADERS) * 16;
      float4 v[PF][2];
      for (int i = 0; i < PF; ++i) {
        if (i < nch) {
          if (staged) { cp_async16(my0 + i * STAGE, g0 + i * KC); cp_async16(my1 + i * STAGE, g1 + i * KC); }
          else { v[i][0] = ld_cg_f4(g0 + i * KC); v[i][1] = ld_cg_f4(g1 + i * KC); }
        }
        cp_async_commit();
      }
      for (int c0 = 0; c0 < nch; c0 += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
          const int c = c0 + i;
          if (c < nch) {
            float4 x0, x1;
            if (staged) {
              cp_async_wait<PF - 1>();
              x0 = *reinterpret_cast<const float4*>(my0 + i * STAGE);
              x1 = *reinterpret_cast<const float4*>(my1 + i * STAGE);
              if (c + PF < nch) { cp_async16(my0 + i * STAGE, g0 + (c + PF) * KC); cp_async16(my1 + i * STAGE, g1 + (c + PF) * KC); }
              cp_async_commit();
            } else {
              x0 = v[i][0]; x1 = v[i][1];
              if (c + PF < nch) { v[i][0] = ld_cg_f4(g0 + (c + PF) * KC); v[i][1] = ld_cg_f4(g1 + (c + PF) * KC); }
            }
            if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
            uint8_t* st = ring + (size_t)slot * SLOT;
            float4 h, l;
            split4(x0, h, l);
            *reinterpret_cast<float4*>(st + so0) = h;
            *reinterpret_cast<float4*>(st + so0 + lo) = l;
            split4(x1, h, l);
            *reinterpret_cast<float4*>(st + so1) = h;
            *reinterpret_cast<float4*>(st + so1 + lo) = l;
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[slot]);
            if (++slot == (uint32_t)nslot) { slot = 0; ++use; }
          }
        }
      }
    } else if (warp == 8) {
      const uint32_t ring_s = smem_u32(ring), w_s = smem_u32(wts);
      const uint32_t idesc = idesc_tf32(128, N);
      for (int c = 0; c < nch; ++c) {
        if (relay > 0) {
          mbar_wait(&ready[slot], use & 1);
        } else {
          mbar_wait(&full[slot], use & 1);
          if (relay == 0) fence_async_smem();
        }
        tc_fence_after();
        if (mma) {
          const uint32_t a0 = ring_s + slot * SLOT, b0 = w_s;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (elect_one()) mma_tf32(tmem, make_desc_sw128(a0 + 32 * j), make_desc_sw128(b0 + 32 * j), idesc, (c | j) ? 1u : 0u);
          if (elect_one()) umma_commit(&empty[slot]);
        } else {
          if (elect_one()) mbar_arrive(&empty[slot]);
        }
        if (++slot == (uint32_t)nslot) { slot = 0; ++use; }
      }
      tc_fence_before();
    } else if (relay > 0 && warp >= 9 && warp < 9 + relay) {
      for (int c = 0; c < nch; ++c) {
        if (c % relay == warp - 9) {  // relay == 3: warps 9, 10, 11 take the chunks round-robin
          mbar_wait(&full[slot], use & 1);
          fence_async_smem();
          if (lane == 0) mbar_arrive(&ready[slot]);
        }
        if (++slot == (uint32_t)nslot) { slot = 0; ++use; }
      }
    }
    // the loaders may run at most nslot chunks ahead, so a CTA-wide barrier per phase mirrors the real kernel
    __syncthreads();
    cc += (uint32_t)nch;
  }
  long long t1 = clock64();
  // drain: the last MMAs must have completed before TMEM goes away
  if (warp == 8 && mma) {
    uint32_t last = (cc - 1) % (uint32_t)nslot, use = (cc - 1) / (uint32_t)nslot;
    mbar_wait(&empty[last], use & 1);
  }
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(64) : "memory");
  }
}

// cycles per fence.proxy.async executed back to back by one warp, alone (mode 0) or after one STS.128 per lane (mode 1)
__global__ void fence_cost(int mode, int n, long long* out) {
  __shared__ float4 buf[32];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (mode) asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};\n" ::"r"(smem_u32(&buf[threadIdx.x & 31])), "f"((float)i) : "memory");
    fence_async_smem();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount, K = 512, reps = 40;
  float* X;
  long long* out;
  cudaMalloc(&X, (size_t)S * 800 * 4);
  cudaMemset(X, 0, (size_t)S * 800 * 4);
  cudaMalloc(&out, nsm * sizeof(long long));
  const size_t smem = 8 * 1024 + MAXSLOT * SLOT + PF * STAGE + 1024 + 1024;
  cudaFuncSetAttribute((const void*)pipe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute((const void*)pipe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 2; ++mode) {
    fence_cost<<<1, 32>>>(mode, 1000, out);
    fence_cost<<<1, 32>>>(mode, 1000, out);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, out, sizeof c, cudaMemcpyDeviceToHost);
    printf("fence.proxy.async (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC)%s: %.1f cycles each\n", mode ? " after an STS.128" : "", c / 1000.0);
  }
  struct Cfg { int generic, staged, mma, nslot, relay, N, wpc; };
  std::vector<Cfg> cfgs;
  for (int generic = 1; generic >= 0; --generic)
    for (int staged = 1; staged >= 0; --staged) {
      cfgs.push_back({generic, staged, 1, 3, 1, 48, 0});   // the shipped configuration (generic, staged) and its variants
      cfgs.push_back({generic, staged, 1, 6, 1, 48, 0});
      cfgs.push_back({generic, staged, 0, 3, 1, 48, 0});   // no MMAs: loader + handshake cost alone
      cfgs.push_back({generic, staged, 1, 3, 0, 48, 0});   // fence in the issuer instead of the relay warp
      cfgs.push_back({generic, staged, 1, 3, 3, 48, 0});   // three relay warps
      cfgs.push_back({generic, staged, 1, 3, -1, 48, 0});  // no proxy fence at all (timing only)
      cfgs.push_back({generic, staged, 1, 3, 1, 16, 0});   // projection-sized MMAs
      if (!staged) {
        cfgs.push_back({generic, 0, 1, 8, 1, 48, 1});      // warp-per-chunk loaders, 8 slots
        cfgs.push_back({generic, 0, 1, 8, 3, 48, 1});      // ... with three relay warps
        cfgs.push_back({generic, 0, 0, 8, 1, 48, 1});      // ... without MMAs
      }
    }
  for (const Cfg& c : cfgs) {
    for (int it = 0; it < 2; ++it) {
      if (c.generic) pipe<true><<<nsm, 384, smem>>>(X, K, reps, c.staged, c.mma, c.nslot, c.relay, c.N, c.wpc, out);
      else pipe<false><<<nsm, 384, smem>>>(X, K, reps, c.staged, c.mma, c.nslot, c.relay, c.N, c.wpc, out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), out, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    printf("%s smem ptrs, %s, mma %d (N=%2d), %d slots, %s: %7.1f cycles per 8 KB chunk (%s)\n",
           c.generic ? "generic" : "shared ", c.wpc ? "WARP-PER-CHUNK " : c.staged ? "cp.async staged" : "LDG registers  ", c.mma, c.N, c.nslot,
           c.relay == 1 ? "relay fence " : c.relay == 3 ? "3 relay warps" : c.relay == 0 ? "issuer fence" : "NO fence    ", (double)mx / (reps * (K / KC)), cudaGetErrorString(e));
  }
  return 0;
}
