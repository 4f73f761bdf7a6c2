/*
 * oracle/lstmp_streams_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's LstmProjectedStreams hot path
 * (dophist/kaldi-lstm @ 284c73cb), op for op, on host row-major matrices:
 *
 *   forward   google/nnet/bd-nnet-lstm-projected-streams.h:222-332  (PropagateFnc)
 *   backward  google/nnet/bd-nnet-lstm-projected-streams.h:334-499  (BackpropagateFnc)
 *   update    google/nnet/bd-nnet-lstm-projected-streams.h:501-512  (Update)
 *   reset     google/nnet/bd-nnet-lstm-projected-streams.h:212-220  (Reset)
 *
 * using the CPU matrix formulas the reference falls through to when no GPU is
 * enabled (google/cudamatrix/cu-matrix.cc:943,1044,1066,1089):
 *
 *   AddMatMat      google/matrix/kaldi-matrix.cc:159-175   (cblas_Xgemm contract)
 *   AddMatDiagVec  google/matrix/kaldi-matrix.cc:447-473
 *   AddMatDotMat   google/matrix/kaldi-matrix.cc:475-497
 *   DiffSigmoid    google/matrix/kaldi-matrix.cc:2561-2576
 *   DiffTanh       google/matrix/kaldi-matrix.cc:2578-2593
 *   AddVecToRows   google/matrix/kaldi-matrix.cc:2596-2609
 *   ApplyFloor/Ceiling  google/matrix/kaldi-matrix.cc:1868-1886
 *   Sigmoid/Tanh   delegate to VectorBase::Sigmoid/Tanh (kaldi-matrix.cc:2464,2552),
 *                  which live in upstream Kaldi and are NOT in the reference tree;
 *                  restated below as the overflow-safe forms Kaldi uses.
 *
 * PARITY PINNED against the reference's own source: the reference ships no tests,
 * fixtures or golden vectors for this path, so oracle/_ref (`make -C oracle ref`)
 * compiles the UNMODIFIED bd-nnet-lstm-projected-streams.h, nnet-lstm-projected.h,
 * nnet-time-shift.h and nnet-loss.cc where they lie under /root/reference, together
 * with the kaldi-matrix.cc / cu-matrix.cc method bodies on the path, against a
 * CPU-computing Kaldi surface (oracle/ref_build/), and tests/test_ref_pin.py checks
 * this restatement against it: bit-for-bit with the same SGEMM, <= 1e-6 otherwise
 * (cfg2, cfg3 both layers, S=1/T=100, reset + carry, clamp saturation, momentum
 * 0 / 0.9).  The .npz fixtures under tests/golden/ were generated from oracle/_ref
 * (tests/golden/make_golden.py).  Independent cross-checks: a torch-autograd
 * restatement of the equations and fp64 finite differences (tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (kaldi-lstm_b200/) never does.
 *
 * Compiled twice: -DORACLE_REAL=float (prefix lstmp_oracle_f32_) and
 * -DORACLE_REAL=double (prefix lstmp_oracle_f64_, error-budget twin).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORACLE_REAL
#define ORACLE_REAL float
#endif
#ifndef ORACLE_PREFIX
#define ORACLE_PREFIX lstmp_oracle_f32_
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(ORACLE_PREFIX, name)

typedef ORACLE_REAL real;

/* Optional external SGEMM (cblas_sgemm ABI), e.g. the OpenBLAS bundled with
 * scipy; set through FN(set_sgemm).  When NULL the blocked loops below run.
 * Only used by the float build (the reference's cblas_Xgemm, kaldi-matrix.cc:172). */
typedef void (*cblas_sgemm_fn)(int order, int transA, int transB, int M, int N, int K,
                               float alpha, const float *A, int lda, const float *B, int ldb,
                               float beta, float *C, int ldc);
static cblas_sgemm_fn g_sgemm = 0;
void FN(set_sgemm)(void *fn) { g_sgemm = (cblas_sgemm_fn)fn; }

/* ------------------------------------------------------------------------- */
/* Matrix primitives (row-major, explicit stride), one per reference op.      */
/* ------------------------------------------------------------------------- */

/* C[M x N] = alpha * op(A) * op(B) + beta * C   -- kaldi-matrix.cc:159-175.
 * tA/tB: 0 = kNoTrans, 1 = kTrans.  op(A) is M x K, op(B) is K x N. */
static void add_mat_mat(real *C, int ldc, int M, int N, real alpha, const real *A, int lda, int tA,
                        const real *B, int ldb, int tB, int K, real beta) {
  if (M == 0 || N == 0) return;
#if defined(ORACLE_IS_FLOAT)
  if (g_sgemm) {
    /* CblasRowMajor=101, CblasNoTrans=111, CblasTrans=112 */
    g_sgemm(101, tA ? 112 : 111, tB ? 112 : 111, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    return;
  }
#endif
  for (int i = 0; i < M; i++) {
    real *c = C + (size_t)i * ldc;
    if (beta == (real)0) {
      for (int j = 0; j < N; j++) c[j] = 0;
    } else if (beta != (real)1) {
      for (int j = 0; j < N; j++) c[j] *= beta;
    }
  }
  if (!tB) {
    /* C[i,:] += alpha * A(i,k) * B[k,:]  (vectorises over j) */
    for (int i = 0; i < M; i++) {
      real *c = C + (size_t)i * ldc;
      for (int k = 0; k < K; k++) {
        real a = alpha * (tA ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k]);
        const real *b = B + (size_t)k * ldb;
        for (int j = 0; j < N; j++) c[j] += a * b[j];
      }
    }
  } else if (!tA) {
    /* C[i,j] += alpha * dot(A[i,:], B[j,:]) */
    for (int i = 0; i < M; i++) {
      const real *a = A + (size_t)i * lda;
      real *c = C + (size_t)i * ldc;
      for (int j = 0; j < N; j++) {
        const real *b = B + (size_t)j * ldb;
        real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int k = 0;
        for (; k + 4 <= K; k += 4) {
          s0 += a[k] * b[k];
          s1 += a[k + 1] * b[k + 1];
          s2 += a[k + 2] * b[k + 2];
          s3 += a[k + 3] * b[k + 3];
        }
        for (; k < K; k++) s0 += a[k] * b[k];
        c[j] += alpha * ((s0 + s1) + (s2 + s3));
      }
    }
  } else {
    for (int i = 0; i < M; i++)
      for (int j = 0; j < N; j++) {
        real s = 0;
        for (int k = 0; k < K; k++) s += A[(size_t)k * lda + i] * B[(size_t)j * ldb + k];
        C[(size_t)i * ldc + j] += alpha * s;
      }
  }
}

/* this = alpha * M * diag(v) + beta * this   -- kaldi-matrix.cc:447-473 (scale first, then +=) */
static void add_mat_diag_vec(real *D, int ldd, int rows, int cols, real alpha, const real *M, int ldm,
                             const real *v, real beta) {
  if (beta != (real)1)
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) D[(size_t)i * ldd + j] *= beta;
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) D[(size_t)i * ldd + j] += alpha * v[j] * M[(size_t)i * ldm + j];
}

/* this = beta * this + alpha * A .* B        -- kaldi-matrix.cc:475-497 */
static void add_mat_dot_mat(real *D, int ldd, int rows, int cols, real alpha, const real *A, int lda,
                            const real *B, int ldb, real beta) {
  for (int i = 0; i < rows; i++) {
    real *d = D + (size_t)i * ldd;
    const real *a = A + (size_t)i * lda, *b = B + (size_t)i * ldb;
    for (int j = 0; j < cols; j++) d[j] = beta * d[j] + alpha * a[j] * b[j];
  }
}

/* Kaldi VectorBase::Sigmoid (upstream kaldi-vector.cc): overflow-safe branches. */
static inline real sigmoid1(real x) {
  /* upstream writes the constants as double literals (`1.0 / (1.0 + Exp(-x))`): with Real = float the exp is
   * expf, the division runs in double and is rounded once -- pinned against oracle/_ref (tests/test_ref_pin.py) */
#if defined(ORACLE_IS_FLOAT)
  if (x > 0.0) return (real)(1.0 / (1.0 + expf(-x)));
  real ex = expf(x);
  return (real)(ex / (ex + 1.0));
#else
  if (x > 0.0) return 1.0 / (1.0 + exp(-x));
  real ex = exp(x);
  return ex / (ex + 1.0);
#endif
}
/* Kaldi VectorBase::Tanh (upstream kaldi-vector.cc); same remark on the double literals. */
static inline real tanh1(real x) {
#if defined(ORACLE_IS_FLOAT)
  if (x > 0.0) {
    real inv_expx = expf(-x);
    return (real)(-1.0 + 2.0 / (1.0 + inv_expx * inv_expx));
  }
  real expx = expf(x);
  return (real)(1.0 - 2.0 / (1.0 + expx * expx));
#else
  if (x > 0.0) {
    real inv_expx = exp(-x);
    return -1.0 + 2.0 / (1.0 + inv_expx * inv_expx);
  }
  real expx = exp(x);
  return 1.0 - 2.0 / (1.0 + expx * expx);
#endif
}
static void sigmoid_mat(real *D, int ldd, int rows, int cols, const real *Sx, int lds) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) D[(size_t)i * ldd + j] = sigmoid1(Sx[(size_t)i * lds + j]);
}
static void tanh_mat(real *D, int ldd, int rows, int cols, const real *Sx, int lds) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) D[(size_t)i * ldd + j] = tanh1(Sx[(size_t)i * lds + j]);
}
/* this = diff .* value .* (1 - value)       -- kaldi-matrix.cc:2561-2576 */
static void diff_sigmoid(real *D, int ldd, int rows, int cols, const real *val, int ldv, const real *diff,
                         int ldf) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      real v = val[(size_t)i * ldv + j];
      D[(size_t)i * ldd + j] = (real)(diff[(size_t)i * ldf + j] * v * (1.0 - v)); /* double literal as in the reference */
    }
}
/* this = diff .* (1 - value^2)              -- kaldi-matrix.cc:2578-2593 */
static void diff_tanh(real *D, int ldd, int rows, int cols, const real *val, int ldv, const real *diff,
                      int ldf) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      real v = val[(size_t)i * ldv + j];
      D[(size_t)i * ldd + j] = (real)(diff[(size_t)i * ldf + j] * (1.0 - (v * v))); /* double literal as in the reference */
    }
}
/* kaldi-matrix.cc:1868-1886 */
static void apply_floor(real *D, int ldd, int rows, int cols, real f) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      real *p = D + (size_t)i * ldd + j;
      *p = (*p < f ? f : *p);
    }
}
static void apply_ceiling(real *D, int ldd, int rows, int cols, real c) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      real *p = D + (size_t)i * ldd + j;
      *p = (*p > c ? c : *p);
    }
}
/* kaldi-matrix.cc:2596-2609 */
static void add_vec_to_rows(real *D, int ldd, int rows, int cols, real alpha, const real *v) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) D[(size_t)i * ldd + j] += alpha * v[j];
}
/* CuVector::AddRowSumMat(alpha, M, beta): v = alpha * colsum(M) + beta * v  (upstream) */
static void add_row_sum_mat(real *v, int cols, real alpha, const real *M, int ldm, int rows, real beta) {
  for (int j = 0; j < cols; j++) {
    real s = 0;
    for (int i = 0; i < rows; i++) s += M[(size_t)i * ldm + j];
    v[j] = alpha * s + beta * v[j];
  }
}
/* CuVector::AddDiagMatMat(alpha, M, kTrans, N, kNoTrans, beta):
 * v[j] = alpha * sum_i M[i,j]*N[i,j] + beta * v[j]   (upstream) */
static void add_diag_mat_mat(real *v, int cols, real alpha, const real *M, int ldm, const real *N, int ldn,
                             int rows, real beta) {
  for (int j = 0; j < cols; j++) {
    real s = 0;
    for (int i = 0; i < rows; i++) s += M[(size_t)i * ldm + j] * N[(size_t)i * ldn + j];
    v[j] = alpha * s + beta * v[j];
  }
}

/* ------------------------------------------------------------------------- */
/* Component state (LPS.h:577-619)                                            */
/* ------------------------------------------------------------------------- */
typedef struct {
  int I, C, R, S;       /* input_dim_, ncell_, nrecur_, nstream_ */
  int W;                /* activation row width 7C+R  (LPS.h:230) */
  real *w_gifo_x, *w_gifo_r, *bias, *p_i, *p_f, *p_o, *w_r_m;
  real *w_gifo_x_corr, *w_gifo_r_corr, *bias_corr, *p_i_corr, *p_f_corr, *p_o_corr, *w_r_m_corr;
  real *prev_state;     /* S x W            (LPS.h:583) */
  real *prop;           /* (T+2)S x W       (LPS.h:616) */
  real *bprop;          /* (T+2)S x W       (LPS.h:619) */
  int T_alloc;
} oracle_t;

static real *zalloc(size_t n) { return (real *)calloc(n ? n : 1, sizeof(real)); }

void *FN(create)(int I, int C, int R, int S) {
  oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
  o->I = I; o->C = C; o->R = R; o->S = S; o->W = 7 * C + R;
  o->w_gifo_x = zalloc((size_t)4 * C * I);  o->w_gifo_x_corr = zalloc((size_t)4 * C * I);
  o->w_gifo_r = zalloc((size_t)4 * C * R);  o->w_gifo_r_corr = zalloc((size_t)4 * C * R);
  o->bias = zalloc((size_t)4 * C);          o->bias_corr = zalloc((size_t)4 * C);
  o->p_i = zalloc(C); o->p_f = zalloc(C); o->p_o = zalloc(C);
  o->p_i_corr = zalloc(C); o->p_f_corr = zalloc(C); o->p_o_corr = zalloc(C);
  o->w_r_m = zalloc((size_t)R * C);         o->w_r_m_corr = zalloc((size_t)R * C);
  o->prev_state = zalloc((size_t)S * o->W);
  return o;
}
void FN(destroy)(void *h) {
  oracle_t *o = (oracle_t *)h;
  if (!o) return;
  free(o->w_gifo_x); free(o->w_gifo_x_corr); free(o->w_gifo_r); free(o->w_gifo_r_corr);
  free(o->bias); free(o->bias_corr); free(o->p_i); free(o->p_f); free(o->p_o);
  free(o->p_i_corr); free(o->p_f_corr); free(o->p_o_corr); free(o->w_r_m); free(o->w_r_m_corr);
  free(o->prev_state); free(o->prop); free(o->bprop); free(o);
}
long FN(num_params)(void *h) {  /* LPS.h:152-160 */
  oracle_t *o = (oracle_t *)h;
  return (long)4 * o->C * o->I + (long)4 * o->C * o->R + 4L * o->C + 3L * o->C + (long)o->R * o->C;
}
/* Flat order of GetParams (LPS.h:162-189):
 * w_gifo_x | w_gifo_r | bias | peephole_i_c | peephole_f_c | peephole_o_c | w_r_m */
static void flat_copy(oracle_t *o, real *flat, int to_flat, int grads) {
  real *srcs[7] = {grads ? o->w_gifo_x_corr : o->w_gifo_x, grads ? o->w_gifo_r_corr : o->w_gifo_r,
                   grads ? o->bias_corr : o->bias,         grads ? o->p_i_corr : o->p_i,
                   grads ? o->p_f_corr : o->p_f,           grads ? o->p_o_corr : o->p_o,
                   grads ? o->w_r_m_corr : o->w_r_m};
  size_t lens[7] = {(size_t)4 * o->C * o->I, (size_t)4 * o->C * o->R, (size_t)4 * o->C, (size_t)o->C,
                    (size_t)o->C, (size_t)o->C, (size_t)o->R * o->C};
  size_t off = 0;
  for (int i = 0; i < 7; i++) {
    if (to_flat) memcpy(flat + off, srcs[i], lens[i] * sizeof(real));
    else memcpy(srcs[i], flat + off, lens[i] * sizeof(real));
    off += lens[i];
  }
}
void FN(set_params)(void *h, const real *flat) { flat_copy((oracle_t *)h, (real *)flat, 0, 0); }
void FN(get_params)(void *h, real *flat) { flat_copy((oracle_t *)h, flat, 1, 0); }
void FN(set_grads)(void *h, const real *flat) { flat_copy((oracle_t *)h, (real *)flat, 0, 1); }
void FN(get_grads)(void *h, real *flat) { flat_copy((oracle_t *)h, flat, 1, 1); }
/* prev_nnet_state_ rows, S x (7C+R) */
void FN(get_state)(void *h, real *dst) {
  oracle_t *o = (oracle_t *)h;
  memcpy(dst, o->prev_state, (size_t)o->S * o->W * sizeof(real));
}
void FN(set_state)(void *h, const real *src) {
  oracle_t *o = (oracle_t *)h;
  memcpy(o->prev_state, src, (size_t)o->S * o->W * sizeof(real));
}
/* propagate_buf_ / backpropagate_buf_ rows [0,(T+2)S), width 7C+R */
const real *FN(prop_buf)(void *h) { return ((oracle_t *)h)->prop; }
const real *FN(bprop_buf)(void *h) { return ((oracle_t *)h)->bprop; }

/* LPS.h:212-220 */
void FN(reset)(void *h, const int *flags, int n) {
  oracle_t *o = (oracle_t *)h;
  for (int s = 0; s < n && s < o->S; s++)
    if (flags[s] == 1) memset(o->prev_state + (size_t)s * o->W, 0, (size_t)o->W * sizeof(real));
}

static void ensure_bufs(oracle_t *o, int T) {
  if (T > o->T_alloc) {
    free(o->prop); free(o->bprop);
    o->prop = (real *)malloc((size_t)(T + 2) * o->S * o->W * sizeof(real));
    o->bprop = (real *)malloc((size_t)(T + 2) * o->S * o->W * sizeof(real));
    o->T_alloc = T;
  }
}

/* LPS.h:222-332.  in: (T*S) x I rows t*S+s; out: (T*S) x R.  Returns 0 / -1 on bad shape. */
int FN(propagate)(void *h, const real *in, int ld_in, real *out, int ld_out, int num_rows) {
  oracle_t *o = (oracle_t *)h;
  const int S = o->S, C = o->C, R = o->R, I = o->I, W = o->W;
  if (num_rows % S != 0) return -1;                                  /* :225 */
  const int T = num_rows / S;                                        /* :226 */
  ensure_bufs(o, T);
  real *P = o->prop;
  memset(P, 0, (size_t)(T + 2) * S * W * sizeof(real));              /* :230 */
  memcpy(P, o->prev_state, (size_t)S * W * sizeof(real));            /* :231 */
#define YROW(t) (P + (size_t)(t) * S * W)
  /* column offsets :234-243 */
  const int oG = 0, oI = C, oF = 2 * C, oO = 3 * C, oC = 4 * C, oH = 5 * C, oM = 6 * C, oR = 7 * C;
  /* :246  YGIFO[1..T] = in * w_gifo_x^T */
  add_mat_mat(YROW(1) + oG, W, T * S, 4 * C, 1, in, ld_in, 0, o->w_gifo_x, I, 1, I, 0);
  /* :259 */
  add_vec_to_rows(YROW(1) + oG, W, T * S, 4 * C, 1, o->bias);
  for (int t = 1; t <= T; t++) {                                     /* :261 */
    real *y = YROW(t), *yp = YROW(t - 1);
    /* :275 */
    add_mat_mat(y + oG, W, S, 4 * C, 1, yp + oR, W, 0, o->w_gifo_r, R, 1, R, 1);
    add_mat_diag_vec(y + oI, W, S, C, 1, yp + oC, W, o->p_i, 1);     /* :278 */
    add_mat_diag_vec(y + oF, W, S, C, 1, yp + oC, W, o->p_f, 1);     /* :281 */
    sigmoid_mat(y + oI, W, S, C, y + oI, W);                         /* :284 */
    sigmoid_mat(y + oF, W, S, C, y + oF, W);                         /* :285 */
    tanh_mat(y + oG, W, S, C, y + oG, W);                            /* :288 */
    add_mat_dot_mat(y + oC, W, S, C, 1, y + oG, W, y + oI, W, 0);    /* :291 */
    add_mat_dot_mat(y + oC, W, S, C, 1, yp + oC, W, y + oF, W, 1);   /* :294 */
    apply_floor(y + oC, W, S, C, (real)-50);                         /* :296 */
    apply_ceiling(y + oC, W, S, C, (real)50);                        /* :297 */
    tanh_mat(y + oH, W, S, C, y + oC, W);                            /* :300 */
    add_mat_diag_vec(y + oO, W, S, C, 1, y + oC, W, o->p_o, 1);      /* :303 */
    sigmoid_mat(y + oO, W, S, C, y + oO, W);                         /* :306 */
    add_mat_dot_mat(y + oM, W, S, C, 1, y + oH, W, y + oO, W, 0);    /* :309 */
    add_mat_mat(y + oR, W, S, R, 1, y + oM, W, 0, o->w_r_m, C, 1, C, 0); /* :312 */
  }
  /* :328 */
  for (int r = 0; r < T * S; r++)
    memcpy(out + (size_t)r * ld_out, YROW(1) + (size_t)r * W + oR, (size_t)R * sizeof(real));
  /* :331 */
  memcpy(o->prev_state, YROW(T), (size_t)S * W * sizeof(real));
#undef YROW
  return 0;
}

/* LPS.h:334-499.  Must follow a propagate() of the same num_rows.  in_diff may be NULL
 * (the reference always computes it; Nnet discards it for the first component). */
int FN(backpropagate)(void *h, const real *in, int ld_in, const real *out_diff, int ld_od, real *in_diff,
                      int ld_id, int num_rows, real momentum) {
  oracle_t *o = (oracle_t *)h;
  const int S = o->S, C = o->C, R = o->R, I = o->I, W = o->W;
  if (num_rows % S != 0) return -1;
  const int T = num_rows / S;                                        /* :338 */
  if (T > o->T_alloc) return -2;
  real *P = o->prop, *B = o->bprop;
  memset(B, 0, (size_t)(T + 2) * S * W * sizeof(real));              /* :352 */
#define YROW(t) (P + (size_t)(t) * S * W)
#define DROW(t) (B + (size_t)(t) * S * W)
  const int oG = 0, oI = C, oF = 2 * C, oO = 3 * C, oC = 4 * C, oH = 5 * C, oM = 6 * C, oR = 7 * C;
  /* :367 */
  for (int r = 0; r < T * S; r++)
    memcpy(DROW(1) + (size_t)r * W + oR, out_diff + (size_t)r * ld_od, (size_t)R * sizeof(real));
  for (int t = T; t >= 1; t--) {                                     /* :369 */
    real *y = YROW(t), *yp = YROW(t - 1), *yn = YROW(t + 1);
    real *d = DROW(t), *dn = DROW(t + 1);
    /* :391  d_r += DGIFO(t+1) * w_gifo_r */
    add_mat_mat(d + oR, W, S, R, 1, dn + oG, W, 0, o->w_gifo_r, R, 0, 4 * C, 1);
    /* :408  d_m = d_r * w_r_m */
    add_mat_mat(d + oM, W, S, C, 1, d + oR, W, 0, o->w_r_m, C, 0, R, 0);
    add_mat_dot_mat(d + oH, W, S, C, 1, d + oM, W, y + oO, W, 0);    /* :411 */
    diff_tanh(d + oH, W, S, C, y + oH, W, d + oH, W);                /* :412 */
    add_mat_dot_mat(d + oO, W, S, C, 1, d + oM, W, y + oH, W, 0);    /* :415 */
    diff_sigmoid(d + oO, W, S, C, y + oO, W, d + oO, W);             /* :416 */
    /* :424  d_c.AddMat(1.0, d_h) */
    for (int s = 0; s < S; s++)
      for (int j = 0; j < C; j++) d[(size_t)s * W + oC + j] += d[(size_t)s * W + oH + j];
    add_mat_dot_mat(d + oC, W, S, C, 1, dn + oC, W, yn + oF, W, 1);  /* :425 */
    add_mat_diag_vec(d + oC, W, S, C, 1, dn + oI, W, o->p_i, 1);     /* :426 */
    add_mat_diag_vec(d + oC, W, S, C, 1, dn + oF, W, o->p_f, 1);     /* :427 */
    add_mat_diag_vec(d + oC, W, S, C, 1, d + oO, W, o->p_o, 1);      /* :428 */
    add_mat_dot_mat(d + oF, W, S, C, 1, d + oC, W, yp + oC, W, 0);   /* :431 */
    diff_sigmoid(d + oF, W, S, C, y + oF, W, d + oF, W);             /* :432 */
    add_mat_dot_mat(d + oI, W, S, C, 1, d + oC, W, y + oG, W, 0);    /* :435 */
    diff_sigmoid(d + oI, W, S, C, y + oI, W, d + oI, W);             /* :436 */
    add_mat_dot_mat(d + oG, W, S, C, 1, d + oC, W, y + oI, W, 0);    /* :439 */
    diff_tanh(d + oG, W, S, C, y + oG, W, d + oG, W);                /* :440 */
  }
  /* :457  in_diff = DGIFO[1..T] * w_gifo_x */
  if (in_diff)
    add_mat_mat(in_diff, ld_id, T * S, I, 1, DROW(1) + oG, W, 0, o->w_gifo_x, I, 0, 4 * C, 0);
  const real mmt = momentum;                                         /* :465 */
  /* :468 */
  add_mat_mat(o->w_gifo_x_corr, I, 4 * C, I, 1, DROW(1) + oG, W, 1, in, ld_in, 0, T * S, mmt);
  /* :471 */
  add_mat_mat(o->w_gifo_r_corr, R, 4 * C, R, 1, DROW(1) + oG, W, 1, YROW(0) + oR, W, 0, T * S, mmt);
  /* :474 */
  add_row_sum_mat(o->bias_corr, 4 * C, 1, DROW(1) + oG, W, T * S, mmt);
  /* :477-484 */
  add_diag_mat_mat(o->p_i_corr, C, 1, DROW(1) + oI, W, YROW(0) + oC, W, T * S, mmt);
  add_diag_mat_mat(o->p_f_corr, C, 1, DROW(1) + oF, W, YROW(0) + oC, W, T * S, mmt);
  add_diag_mat_mat(o->p_o_corr, C, 1, DROW(1) + oO, W, YROW(1) + oC, W, T * S, mmt);
  /* :486 */
  add_mat_mat(o->w_r_m_corr, C, R, C, 1, DROW(1) + oR, W, 1, YROW(1) + oM, W, 0, T * S, mmt);
#undef YROW
#undef DROW
  return 0;
}

/* LPS.h:501-512 (args unused there too) */
void FN(update)(void *h, real lr) {
  oracle_t *o = (oracle_t *)h;
  size_t n;
  n = (size_t)4 * o->C * o->I; for (size_t i = 0; i < n; i++) o->w_gifo_x[i] += -lr * o->w_gifo_x_corr[i];
  n = (size_t)4 * o->C * o->R; for (size_t i = 0; i < n; i++) o->w_gifo_r[i] += -lr * o->w_gifo_r_corr[i];
  n = (size_t)4 * o->C;        for (size_t i = 0; i < n; i++) o->bias[i] += -lr * o->bias_corr[i];
  n = (size_t)o->C;            for (size_t i = 0; i < n; i++) o->p_i[i] += -lr * o->p_i_corr[i];
  for (size_t i = 0; i < n; i++) o->p_f[i] += -lr * o->p_f_corr[i];
  for (size_t i = 0; i < n; i++) o->p_o[i] += -lr * o->p_o_corr[i];
  n = (size_t)o->R * o->C;     for (size_t i = 0; i < n; i++) o->w_r_m[i] += -lr * o->w_r_m_corr[i];
}

/* standard/nnet/nnet-lstm-projected.h:482-493 -- the single-stream variant clips every
 * gradient element to +-50 before the SGD step; exposed so S=1 parity with that file can
 * be exercised.  Not used by the streams path. */
void FN(clip_grads)(void *h, real max_grad) {
  oracle_t *o = (oracle_t *)h;
  real *g[7] = {o->w_gifo_x_corr, o->w_gifo_r_corr, o->bias_corr, o->p_i_corr, o->p_f_corr, o->p_o_corr,
                o->w_r_m_corr};
  size_t lens[7] = {(size_t)4 * o->C * o->I, (size_t)4 * o->C * o->R, (size_t)4 * o->C, (size_t)o->C,
                    (size_t)o->C, (size_t)o->C, (size_t)o->R * o->C};
  for (int k = 0; k < 7; k++)
    for (size_t i = 0; i < lens[k]; i++) {
      if (g[k][i] < -max_grad) g[k][i] = -max_grad;
      if (g[k][i] > max_grad) g[k][i] = max_grad;
    }
}
