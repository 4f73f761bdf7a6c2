// Persistent per-chunk forward / BPTT kernels for LstmProjectedStreams on sm_100a.
//
// Replaces the reference's per-timestep launch sequence
//   forward  google/nnet/bd-nnet-lstm-projected-streams.h:261-325 (15 launches / step)
//   backward google/nnet/bd-nnet-lstm-projected-streams.h:369-454 (17 launches / step)
// with ONE cooperative launch per chunk and direction.  Every CTA keeps its slice of
// W_gifo_r / W_r_m in shared memory for the whole chunk (TMA bulk-staged once), owns a
// slice of cells for all streams of its stream group, and exchanges r_t / m_t / d_r with
// the other CTAs of the group through L2 + a group barrier.
//
// Design notes live in DESIGN.md ("Kernels").
#include "lstmp_common.cuh"
#include "lstmp_kernels.h"

namespace lstmp {

// =========================================================================================
// Forward
// =========================================================================================
__global__ void __launch_bounds__(kThreads, 1) lstmp_fwd_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t mbar;

  const int tid = threadIdx.x;
  const int grp = blockIdx.x / p.d.ctas_per_group;
  const int j = blockIdx.x - grp * p.d.ctas_per_group;
  const int C = p.C, R = p.R, S = p.S, T = p.T;
  const int Sg = p.d.Sg;
  const int s_base = grp * Sg;
  const int c0 = j * p.d.cpc;
  const int nc = max(0, min(p.d.cpc, C - c0));  // my cells
  const int r0 = j * p.d.rpc;
  const int nr = max(0, min(p.d.rpc, R - r0));  // my projection outputs

  float* wr = smem + p.off_wr;      // [4*nc][ldwr]   rows: gate*nc + cl  (gate order g,i,f,o: LPS.h:234-243)
  float* wm = smem + p.off_wm;      // [nr][ldwm]
  float* xbuf = smem + p.off_xbuf;  // 2 x [Sg][ldx]
  float* red = smem + p.off_red;    // [Sg][ldred]
  float* cprev = smem + p.off_cprev;  // [Sg*nc]  c_{t-1} of my cells
  float* peep = smem + p.off_peep;    // [3][cpc]
  SkinnyPlan* plan_g = reinterpret_cast<SkinnyPlan*>(smem + p.off_plan);  // gates:      [Sg x R] * [4nc x R]^T
  SkinnyPlan* plan_p = plan_g + 1;                                         // projection: [Sg x C] * [nr x C]^T
  skinny_make_plan(plan_g, Sg, 4 * nc, R, p.d.fwd_xcap);
  skinny_make_plan(plan_p, Sg, nr, C, p.d.fwd_xcap);

  // ---- stage the stationary weight slices with TMA bulk copies (once per chunk) -----------
  if (tid == 0) {
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t stage_bytes = (uint32_t)((4 * nc * R + nr * C) * sizeof(float));
  if (tid == 0 && stage_bytes) mbar_arrive_expect_tx(&mbar, stage_bytes);
  for (int row = tid; row < 4 * nc; row += kThreads) {
    int gate = row / nc, cl = row - gate * nc;
    tma_bulk_g2s(wr + (size_t)row * p.ldwr, p.w_gifo_r + (size_t)(gate * C + c0 + cl) * R,
                 (uint32_t)(R * sizeof(float)), &mbar);
  }
  for (int row = tid; row < nr; row += kThreads)
    tma_bulk_g2s(wm + (size_t)row * p.ldwm, p.w_r_m + (size_t)(r0 + row) * C, (uint32_t)(C * sizeof(float)),
                 &mbar);

  // ---- history: c_0 of my cells -> smem and cbuf block 0; r_0 of my columns -> rbuf block 0
  // (LPS.h:231 propagate_buf_.RowRange(0,S).CopyFromMat(prev_nnet_state_))
  for (int idx = tid; idx < Sg * nc; idx += kThreads) {
    int s = idx / nc, cl = idx - s * nc;
    float v = p.state_c[(size_t)(s_base + s) * C + c0 + cl];
    cprev[idx] = v;
    p.cbuf[(size_t)(s_base + s) * C + c0 + cl] = v;
  }
  for (int idx = tid; idx < Sg * nr; idx += kThreads) {
    int s = idx / nr, n = idx - s * nr;
    p.rbuf[(size_t)(s_base + s) * R + r0 + n] = p.state_r[(size_t)(s_base + s) * R + r0 + n];
  }
  for (int cl = tid; cl < nc; cl += kThreads) {
    peep[cl] = p.p_i[c0 + cl];
    peep[p.d.cpc + cl] = p.p_f[c0 + cl];
    peep[2 * p.d.cpc + cl] = p.p_o[c0 + cl];
  }
  if (stage_bytes) mbar_wait(&mbar, 0);
  __syncthreads();

  GroupBarrier gb;
  gb.init(p.bar + grp * kBarStride, p.bar_base[grp], (unsigned)p.d.ctas_per_group, (p.d.dbg & 1) != 0);
  stamp_begin((p.d.dbg & 4) && blockIdx.x == 0);
  __syncthreads();
  stamp(1);

  for (int tt = 0; tt < T; ++tt) {
    stamp(10);
    // ================= phase 1: gates + cell update for my cells ==========================
    if (nc > 0) {
      // prefetch this thread's first element's x-part pre-activations (input GEMM + bias)
      float xg = 0.f, xi = 0.f, xf = 0.f, xo = 0.f;
      if (tid < Sg * nc) {
        int s = tid / nc, cl = tid - s * nc;
        const float* gp = p.gifo + (size_t)(tt * S + s_base + s) * (4 * C) + c0 + cl;
        xg = gp[0];
        xi = gp[C];
        xf = gp[2 * C];
        xo = gp[3 * C];
      }
      // r_{t-1}: carried state for the first frame of the chunk, else rbuf block tt
      const float* X = (tt == 0) ? p.state_r + (size_t)s_base * R : p.rbuf + ((size_t)tt * S + s_base) * R;
      // gifo(t) += r(t-1) * W_gifo_r^T                                       (LPS.h:275)
      if (!(p.d.dbg & 2)) skinny_gemm(plan_g, X, (size_t)R, wr, p.ldwr, xbuf, red, p.ldred);
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
        const float* rr = red + s * p.ldred + cl;
        float cp = cprev[idx];
        float pi = peep[cl], pf = peep[p.d.cpc + cl], po = peep[2 * p.d.cpc + cl];
        float ai = rr[nc] + xi + cp * pi;        // :278  i += c(t-1) .* peephole_i_c
        float af = rr[2 * nc] + xf + cp * pf;    // :281
        float gi = sigmoidf_fast(ai);            // :284
        float gf = sigmoidf_fast(af);            // :285
        float gg = tanhf_fast(rr[0] + xg);       // :288
        float c = gg * gi + cp * gf;             // :291-294
        c = fminf(fmaxf(c, -kCellClip), kCellClip);  // :296-297
        float h = tanhf_fast(c);                 // :300
        float ao = rr[3 * nc] + xo + c * po;     // :303  (uses c(t), post-clip)
        float go = sigmoidf_fast(ao);            // :306
        float m = h * go;                        // :309
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
      if (!(p.d.dbg & 2)) skinny_gemm(plan_p, p.mbuf + ((size_t)tt * S + s_base) * C, (size_t)C, wm, p.ldwm, xbuf, red, p.ldred);
      for (int idx = tid; idx < Sg * nr; idx += kThreads) {
        int s = idx / nr, n = idx - s * nr;
        float v = red[s * p.ldred + n];
        size_t row = (size_t)tt * S + s_base + s;
        p.rbuf[(row + S) * R + r0 + n] = v;
        p.out[row * p.ld_out + r0 + n] = v;                                  // :328
        if (tt == T - 1) p.state_r[(size_t)(s_base + s) * R + r0 + n] = v;   // :331
      }
    }
    stamp(30);
    if (tt + 1 < T) gb.sync();
    stamp(31);
  }
  // prev_nnet_state_ <- last frame (LPS.h:331): c part
  for (int idx = tid; idx < Sg * nc; idx += kThreads) {
    int s = idx / nc, cl = idx - s * nc;
    p.state_c[(size_t)(s_base + s) * C + c0 + cl] = cprev[idx];
  }
  stamp_flush(p.d.dbg_stamps);
}

int fwd_barriers(int T) { return 2 * T - 1; }  // barrier epochs consumed by one launch

// The all-gather buffer takes whatever shared memory is left beside the weights; it must hold at least
// two [Sg x 128] ring slots.  The device-side plan decides between a resident panel and a ring.
static bool xbuf_capacity(size_t used_floats, size_t limit_floats, int Sg, int* cap) {
  const size_t avail = limit_floats > used_floats ? limit_floats - used_floats : 0;
  if (avail < (size_t)2 * Sg * (128 + kXbufPadMax)) return false;
  *cap = (int)(avail & ~size_t(3));
  return true;
}

int static_smem_reserve() { return kStaticSmemReserve; }

size_t fwd_smem_floats(int C, int R, Decomp& d, FwdParams* p, size_t limit_floats) {
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t o = off;
    off += (n + 3) & ~size_t(3);
    return (int)o;
  };
  // row stride = K + 16 floats: for K % 32 == 0 two consecutive rows sit in opposite halves of the 32 banks, so a
  // warp's LDS.128 over (2 rows x 4 quads) is one wavefront (with K + 4 it was two)
  p->ldwr = R + 16;
  p->off_wr = take((size_t)4 * d.cpc * p->ldwr);
  p->ldwm = C + 16;
  p->off_wm = take((size_t)d.rpc * p->ldwm);
  int ldred = 4 * d.cpc > d.rpc ? 4 * d.cpc : d.rpc;
  p->ldred = ldred | 1;
  p->off_red = take((size_t)d.Sg * p->ldred);
  p->off_cprev = take((size_t)d.Sg * d.cpc);
  p->off_peep = take((size_t)3 * d.cpc);
  p->off_plan = take(2 * (sizeof(SkinnyPlan) / sizeof(float)) + 4);
  if (!xbuf_capacity(off + 16, limit_floats, d.Sg, &d.fwd_xcap)) return (size_t)1 << 40;
  // no need for more than the largest full panel
  size_t want = (size_t)d.Sg * ((((C > R ? C : R) + 31) & ~31) + kXbufPadMax);
  const size_t ring_min = (size_t)2 * d.Sg * (128 + kXbufPadMax);  // what the device-side ring plan assumes
  if (want < ring_min) want = ring_min;
  if ((size_t)d.fwd_xcap > want) d.fwd_xcap = (int)want;
  p->off_xbuf = take((size_t)d.fwd_xcap);
  return off;
}

cudaError_t launch_fwd(const FwdParams& p, size_t smem_bytes, cudaStream_t stream) {
  void* args[] = {(void*)&p};
  dim3 grid(p.d.ngroups * p.d.ctas_per_group), block(kThreads);
  return cudaLaunchCooperativeKernel((const void*)lstmp_fwd_kernel, grid, block, args, smem_bytes, stream);
}

// =========================================================================================
// Backward
// =========================================================================================

// P[s][k] = sum_{n<Nc} dT[n][s] * W[n][k]   (partial d_r over my gate rows; LPS.h:391 restricted
// to the rows of W_gifo_r this CTA holds).  8 streams x 8 k register tile per thread (two lane-contiguous
// quads of k, R/2 apart), result to global.  The loop is bound by LDS.128 issue (one W request is shared by
// 8 streams here, by 8 streams x 4 k before: the 8x8 tile halves the requests per FMA); the weight pairs
// (k, k+1) come straight from the 128-bit loads, the stream value is duplicated into an FFMA2 operand pair.
__device__ __forceinline__ void outer_gemm(const float* __restrict__ dT, int ldd, const float* __restrict__ Ws,
                                           int ldw, int Nc, int Sg, int R, float* __restrict__ P) {
  const int nkq = R >> 2;           // quads along k
  const int half = (nkq + 1) >> 1;  // a thread owns quads kq and kq + half (the second one may not exist: R % 8 == 4)
  const int nst = ceil_div(Sg, 8);
  const int tiles = half * nst;
  for (int tile = threadIdx.x; tile < tiles; tile += kThreads) {
    const int kq = tile % half, st = tile / half;
    f32x2 acc2[8][4];  // [stream][k pair]: pairs 0,1 = quad kq, pairs 2,3 = quad kq + half
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc2[i][c] = 0ull;
    const bool has2 = kq + half < nkq;
    const float* wp0 = Ws + 4 * kq;
    const float* wp1 = has2 ? Ws + 4 * (kq + half) : wp0;
    const float* dp = dT + 8 * st;
#pragma unroll 2
    for (int n = 0; n < Nc; ++n) {
      const float4 w0 = *reinterpret_cast<const float4*>(wp0 + (size_t)n * ldw);
      const float4 w1 = *reinterpret_cast<const float4*>(wp1 + (size_t)n * ldw);
      const float4 d0 = *reinterpret_cast<const float4*>(dp + (size_t)n * ldd);
      const float4 d1 = *reinterpret_cast<const float4*>(dp + (size_t)n * ldd + 4);
      const f32x2 wv[4] = {pack2(w0.x, w0.y), pack2(w0.z, w0.w), pack2(w1.x, w1.y), pack2(w1.z, w1.w)};
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const f32x2 dd = pack2(dv[i], dv[i]);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc2[i][c] = ffma2(dd, wv[c], acc2[i][c]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int sr = 8 * st + i;
      if (sr < Sg) {
        float4 o0, o1;
        unpack2(acc2[i][0], o0.x, o0.y);
        unpack2(acc2[i][1], o0.z, o0.w);
        unpack2(acc2[i][2], o1.x, o1.y);
        unpack2(acc2[i][3], o1.z, o1.w);
        *reinterpret_cast<float4*>(P + (size_t)sr * R + 4 * kq) = o0;
        if (has2) *reinterpret_cast<float4*>(P + (size_t)sr * R + 4 * (kq + half)) = o1;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) lstmp_bwd_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t mbar;

  const int tid = threadIdx.x;
  const int grp = blockIdx.x / p.d.ctas_per_group;
  const int j = blockIdx.x - grp * p.d.ctas_per_group;
  const int C = p.C, R = p.R, S = p.S, T = p.T;
  const int Sg = p.d.Sg;
  const int s_base = grp * Sg;
  const int c0 = j * p.d.cpc;
  const int nc = max(0, min(p.d.cpc, C - c0));
  const int nper = Sg * R;  // floats in the group's d_r block
  const int e0 = min(nper, j * p.d.piece);
  const int e1 = min(nper, e0 + p.d.piece);  // my reduce-scatter slice [e0,e1)

  float* wr = smem + p.off_wr;      // [4*nc][ldwr]
  float* wmt = smem + p.off_wmt;    // [nc][ldwmt]   W_r_m[:, my cells] transposed
  float* xbuf = smem + p.off_xbuf;
  float* red = smem + p.off_red;    // [Sg][ldred]
  float* dgn = smem + p.off_dgn;    // [4*nc][ldd]  DGIFO(t+1) of my cells, transposed (stream-contiguous)
  float* dcn = smem + p.off_dcn;    // [Sg*nc]      d_c(t+1)
  float* acc7 = smem + p.off_acc7;  // [Sg*nc][7]   running sums for bias / peephole gradients
  float* peep = smem + p.off_peep;
  SkinnyPlan* plan_m = reinterpret_cast<SkinnyPlan*>(smem + p.off_plan);  // d_m: [Sg x R] * [nc x R]^T
  skinny_make_plan(plan_m, Sg, nc, R, p.d.bwd_xcap);

  if (tid == 0) {
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t stage_bytes = (uint32_t)(4 * nc * R * sizeof(float));
  if (tid == 0 && stage_bytes) mbar_arrive_expect_tx(&mbar, stage_bytes);
  for (int row = tid; row < 4 * nc; row += kThreads) {
    int gate = row / nc, cl = row - gate * nc;
    tma_bulk_g2s(wr + (size_t)row * p.ldwr, p.w_gifo_r + (size_t)(gate * C + c0 + cl) * R,
                 (uint32_t)(R * sizeof(float)), &mbar);
  }
  // W_r_m[k][c0+cl] -> wmt[cl][k]
  for (int idx = tid; idx < R * nc; idx += kThreads) {
    int k = idx / nc, cl = idx - k * nc;
    wmt[(size_t)cl * p.ldwmt + k] = p.w_r_m[(size_t)k * C + c0 + cl];
  }
  for (int idx = tid; idx < 4 * p.d.cpc * p.ldd; idx += kThreads) dgn[idx] = 0.f;  // row-block T+1 is zero (:352)
  for (int idx = tid; idx < Sg * p.d.cpc; idx += kThreads) dcn[idx] = 0.f;
  for (int idx = tid; idx < Sg * p.d.cpc * 7; idx += kThreads) acc7[idx] = 0.f;
  for (int cl = tid; cl < nc; cl += kThreads) {
    peep[cl] = p.p_i[c0 + cl];
    peep[p.d.cpc + cl] = p.p_f[c0 + cl];
    peep[2 * p.d.cpc + cl] = p.p_o[c0 + cl];
  }
  if (stage_bytes) mbar_wait(&mbar, 0);
  __syncthreads();

  GroupBarrier gb;
  gb.init(p.bar + grp * kBarStride, p.bar_base[grp], (unsigned)p.d.ctas_per_group, (p.d.dbg & 1) != 0);
  stamp_begin((p.d.dbg & 8) && blockIdx.x == 0);
  __syncthreads();
  stamp(2);
  float* my_scratch = p.scratch + ((size_t)grp * p.d.ctas_per_group + j) * nper;
  const float* grp_scratch = p.scratch + (size_t)grp * p.d.ctas_per_group * nper;
  // number of CTAs of the group that own cells (the only ones that write partials)
  const int nprod = ceil_div(C, p.d.cpc) < p.d.ctas_per_group ? ceil_div(C, p.d.cpc) : p.d.ctas_per_group;

  for (int tt = T - 1; tt >= 0; --tt) {
    const bool have_next = (tt + 1 < T);
    // ============ phase A: partial d_r(t) = DGIFO(t+1)[:, my rows] * W_gifo_r[my rows, :]   (:391)
    stamp(40);
    if (have_next) {
      if (nc > 0 && !(p.d.dbg & 2)) outer_gemm(dgn, p.ldd, wr, p.ldwr, 4 * nc, Sg, R, my_scratch);
      stamp(41);
      gb.sync();
      stamp(42);
    }
    // ============ phase A2: reduce-scatter.  d_r(t)[e0:e1) = out_diff(t) + sum of partials    (:367,:391)
    {
      const int nq = (e1 - e0) >> 2;  // float4 columns in my slice
      const int ncw = nq < kThreads ? nq : kThreads;
      for (int cb = 0; cb < nq; cb += kThreads) {
        const int w = (nq - cb) < kThreads ? (nq - cb) : kThreads;  // columns in this batch
        const int npg = have_next ? max(1, kThreads / w) : 1;
        const int col = tid % w, pg = tid / w;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool on = (pg < npg);
        if (on && have_next) {
          // batches of 12 independent L2 loads, then the (fixed-order) adds: one load per add would serialise the
          // ~700-cycle L2 round trips (the stamps showed 7-9 k cycles for this phase)
          const float* src = grp_scratch + e0 + 4 * (cb + col);
          for (int q = pg; q < nprod; q += 12 * npg) {
            float4 v[12];
#pragma unroll
            for (int u = 0; u < 12; ++u) {
              const int qq = q + u * npg;
              v[u] = qq < nprod ? ld_cg_f4(src + (size_t)qq * nper) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 12; ++u) {
              a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w;
            }
          }
        }
        float4* rb = reinterpret_cast<float4*>(xbuf);
        if (npg > 1) {
          __syncthreads();
          if (on) rb[pg * w + col] = a;
          __syncthreads();
        }
        if (pg == 0) {
          for (int q = 1; q < npg; ++q) {
            float4 v = rb[q * w + col];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
          }
          const int e = e0 + 4 * (cb + col);
          const int s = e / R, k = e - s * R;
          const size_t row = (size_t)tt * S + s_base + s;
          const float* od = p.out_diff + row * p.ld_od + k;
          a.x += od[0]; a.y += od[1]; a.z += od[2]; a.w += od[3];
          *reinterpret_cast<float4*>(p.dr + row * R + k) = a;
        }
      }
      (void)ncw;
    }
    stamp(43);
    gb.sync();
    stamp(44);
    // ============ phase B: d_m = d_r * W_r_m (my cells) (:408) + gate derivatives (:411-440)
    if (nc > 0) {
      // prefetch this thread's first element's activations while d_r is gathered and contracted
      float yg = 0.f, yi = 0.f, yf = 0.f, yo = 0.f, yc = 0.f, ycp = 0.f, yh = 0.f, yfn = 0.f;
      if (tid < Sg * nc) {
        int s = tid / nc, cl = tid - s * nc;
        size_t row = (size_t)tt * S + s_base + s;
        const float* gp = p.gifo + row * (4 * C) + c0 + cl;
        yg = gp[0]; yi = gp[C]; yf = gp[2 * C]; yo = gp[3 * C];
        yc = p.cbuf[(row + S) * C + c0 + cl];
        ycp = p.cbuf[row * C + c0 + cl];
        yh = p.hbuf[row * C + c0 + cl];
        yfn = have_next ? gp[(size_t)S * 4 * C + 2 * C] : 0.f;
      }
      if (!(p.d.dbg & 2)) skinny_gemm(plan_m, p.dr + ((size_t)tt * S + s_base) * R, (size_t)R, wmt, p.ldwmt, xbuf, red, p.ldred);
      for (int idx = tid; idx < Sg * nc; idx += kThreads) {
        int s = idx / nc, cl = idx - s * nc;
        size_t row = (size_t)tt * S + s_base + s;
        if (idx >= kThreads) {
          const float* gp = p.gifo + row * (4 * C) + c0 + cl;
          yg = gp[0]; yi = gp[C]; yf = gp[2 * C]; yo = gp[3 * C];
          yc = p.cbuf[(row + S) * C + c0 + cl];
          ycp = p.cbuf[row * C + c0 + cl];  // c(t-1): block tt
          yh = p.hbuf[row * C + c0 + cl];
          yfn = have_next ? gp[(size_t)S * 4 * C + 2 * C] : 0.f;  // f(t+1)
        }
        float pi = peep[cl], pf = peep[p.d.cpc + cl], po = peep[2 * p.d.cpc + cl];
        float d_m = red[s * p.ldred + cl];
        float d_h = (d_m * yo) * (1.0f - yh * yh);            // :411-412
        float d_o = (d_m * yh) * yo * (1.0f - yo);            // :415-416
        float d_c = d_h;                                      // :424
        d_c += dcn[idx] * yfn;                                // :425
        d_c += dgn[(size_t)(1 * nc + cl) * p.ldd + s] * pi;   // :426
        d_c += dgn[(size_t)(2 * nc + cl) * p.ldd + s] * pf;   // :427
        d_c += d_o * po;                                      // :428
        float d_f = (d_c * ycp) * yf * (1.0f - yf);           // :431-432
        float d_i = (d_c * yg) * yi * (1.0f - yi);            // :435-436
        float d_g = (d_c * yi) * (1.0f - yg * yg);            // :439-440
        float* dp = p.dgifo + row * (4 * C) + c0 + cl;
        dp[0] = d_g;
        dp[C] = d_i;
        dp[2 * C] = d_f;
        dp[3 * C] = d_o;
        dgn[(size_t)(0 * nc + cl) * p.ldd + s] = d_g;
        dgn[(size_t)(1 * nc + cl) * p.ldd + s] = d_i;
        dgn[(size_t)(2 * nc + cl) * p.ldd + s] = d_f;
        dgn[(size_t)(3 * nc + cl) * p.ldd + s] = d_o;
        dcn[idx] = d_c;
        float* a7 = acc7 + (size_t)idx * 7;
        a7[0] += d_g;          // bias_corr_ column sums (:474)
        a7[1] += d_i;
        a7[2] += d_f;
        a7[3] += d_o;
        a7[4] += d_i * ycp;    // peephole_i_c_corr_  DI(t) .* C(t-1)  (:477)
        a7[5] += d_f * ycp;    // peephole_f_c_corr_                  (:480)
        a7[6] += d_o * yc;     // peephole_o_c_corr_  DO(t) .* C(t)    (:483)
      }
      __syncthreads();  // dgn complete before phase A of the next (earlier) frame reads it
      stamp(45);
    }
  }
  // per-group partial bias / peephole gradients: sum my cells over my streams (fixed order)
  __syncthreads();
  float* sg = p.small_grads + (size_t)grp * 7 * C;
  for (int q = tid; q < nc * 7; q += kThreads) {
    int cl = q / 7, w = q - cl * 7;
    float a = 0.f;
    for (int s = 0; s < Sg; ++s) a += acc7[(size_t)(s * nc + cl) * 7 + w];
    sg[(size_t)w * C + c0 + cl] = a;
  }
  stamp_flush(p.d.dbg_stamps);
}

int bwd_barriers(int T) { return 2 * T - 1; }

size_t bwd_smem_floats(int C, int R, Decomp& d, BwdParams* p, size_t limit_floats) {
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t o = off;
    off += (n + 3) & ~size_t(3);
    return (int)o;
  };
  (void)C;
  p->ldwr = R + 16;
  p->off_wr = take((size_t)4 * d.cpc * p->ldwr);
  p->ldwmt = R + 16;
  p->off_wmt = take((size_t)d.cpc * p->ldwmt);
  p->ldred = d.cpc | 1;
  p->off_red = take((size_t)d.Sg * p->ldred);
  p->ldd = (d.Sg + 7) & ~7;
  p->off_dgn = take((size_t)4 * d.cpc * p->ldd);
  p->off_dcn = take((size_t)d.Sg * d.cpc);
  p->off_acc7 = take((size_t)d.Sg * d.cpc * 7);
  p->off_peep = take((size_t)3 * d.cpc);
  p->off_plan = take((sizeof(SkinnyPlan) / sizeof(float)) + 4);
  if (!xbuf_capacity(off + 16 + 4 * kThreads, limit_floats, d.Sg, &d.bwd_xcap)) return (size_t)1 << 40;
  size_t want = (size_t)d.Sg * (((R + 31) & ~31) + kXbufPadMax);
  const size_t ring_min = (size_t)2 * d.Sg * (128 + kXbufPadMax);
  if (want < ring_min) want = ring_min;
  if ((size_t)d.bwd_xcap > want) d.bwd_xcap = (int)want;
  size_t xb = (size_t)d.bwd_xcap;
  if (xb < (size_t)4 * kThreads) xb = (size_t)4 * kThreads;  // also the reduce-scatter exchange buffer
  p->off_xbuf = take(xb);
  return off;
}

cudaError_t launch_bwd(const BwdParams& p, size_t smem_bytes, cudaStream_t stream) {
  void* args[] = {(void*)&p};
  dim3 grid(p.d.ngroups * p.d.ctas_per_group), block(kThreads);
  return cudaLaunchCooperativeKernel((const void*)lstmp_bwd_kernel, grid, block, args, smem_bytes, stream);
}

cudaError_t set_kernel_smem_limits(size_t fwd_bytes, size_t bwd_bytes) {
  // The attribute is per device and per function; several engines (e.g. stacked layers) may
  // share this process, so only ever raise it.
  static size_t cur_fwd[64] = {0}, cur_bwd[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  dev &= 63;
  if (fwd_bytes > cur_fwd[dev]) {
    e = cudaFuncSetAttribute((const void*)lstmp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)fwd_bytes);
    if (e != cudaSuccess) return e;
    cur_fwd[dev] = fwd_bytes;
  }
  if (bwd_bytes > cur_bwd[dev]) {
    e = cudaFuncSetAttribute((const void*)lstmp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)bwd_bytes);
    if (e != cudaSuccess) return e;
    cur_bwd[dev] = bwd_bytes;
  }
  return cudaSuccess;
}

}  // namespace lstmp
