// tcgen05 / TMEM / UMMA-descriptor helpers shared by the tensor-core GEMMs (lstmp_gemm_hl.cu, lstmp_gemm_tc.cu) and
// the tensor-core time loops (lstmp_recurrent_tma.cu).  sm_100a only.
//
// Shared-memory operand layout used everywhere: SWIZZLE_128B K-major tiles -- rows of 128 bytes along K (32 tf32 or
// 64 bf16), 8-row groups of 1024 bytes, tile bases 1024-byte aligned, the 16-byte unit index of a row XORed with
// (row & 7); see sw128_off / make_desc_sw128 below.  (make_desc, the no-swizzle form, is kept for the microbenchmarks
// under tools/.)
#pragma once
#include "lstmp_common.cuh"

namespace lstmp {
namespace tc {

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
  // layout_type SWIZZLE_NONE=0 [61,64)
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Two 16-column loads in flight, one wait (the accumulator read-out of the time loops adds the hi / lo column halves).
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr_a, uint32_t taddr_b, float* a, float* b) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr_a));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr_b));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(q[i]);
  }
}

// SWIZZLE_128B K-major operand tile (the layout every tcgen05 kernel here uses since the no-swizzle "interleave"
// layout turned out to feed the tensor core at only ~30 B/clk -- ~200-260 cycles per M=128,K=8 tf32 MMA whatever N):
// rows of 128 bytes (32 fp32 along K), 8-row groups of 1024 bytes, tile base 1024-byte aligned; the 16-byte chunk
// index kc of a row is XORed with (row & 7) (cute Swizzle<3,4,3> on the byte offset).  One K = 8 MMA covers chunks
// 2j, 2j+1: its descriptor start address is tile + 32*j bytes (the hardware applies the XOR to the address bits).
__host__ __device__ constexpr uint32_t sw128_off(int row, int kc) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((kc ^ row) & 7) << 4));
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  // start>>4 [0,14) | LBO>>4 = 1 (unused for swizzled K-major) [16,30) | SBO>>4 = 64 (1024 B between 8-row groups)
  // [32,46) | version=1 [46,48) | layout_type SWIZZLE_128B=2 [61,64)      (cute::UMMA::SmemDescriptor)
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// How tcgen05.mma must be ISSUED (tools/umma_rate_bench.cu, measured on B200): inside a lane-0 branch with operands in
// ordinary registers ptxas wraps every UTCHMMA in R2UR moves + a BRA.U.ANY "waterfall" loop and the warp issues one
// MMA per ~180-230 cycles whatever its size; with the whole warp executing the (unrolled) burst on warp-uniform
// operands and only the instruction predicated by elect.sync, every MMA of the burst gets its own uniform registers
// and the issue rate is the tensor-pipe floor (N/2 cycles, >= 40).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, both operands K-major:
// D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2, a_major [15]=0, b_major [16]=0, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace tc
}  // namespace lstmp
