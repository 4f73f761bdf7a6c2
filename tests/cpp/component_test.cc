// C++ test of the Kaldi-side host mirror (kaldi-lstm_b200/kaldi/b200-lstm-projected-streams.h) against the CPU
// oracle (tests may link the oracle; the product never does).  Needs a B200.  Prints "PASS" / returns 0.
#include <dlfcn.h>

#include <cstdio>
#include <random>
#include <sstream>

#include "b200-lstm-projected-streams.h"

using namespace kaldi;
using namespace kaldi::nnet1;

struct Oracle {  // oracle/lstmp_streams_oracle.c, fp32 entry points
  void* lib;
  void* (*create)(int, int, int, int);
  void (*destroy)(void*);
  void (*set_params)(void*, const float*);
  void (*get_params)(void*, float*);
  void (*get_grads)(void*, float*);
  void (*reset)(void*, const int*, int);
  int (*propagate)(void*, const float*, int, float*, int, int);
  int (*backpropagate)(void*, const float*, int, const float*, int, float*, int, int, float);
  void (*update)(void*, float);
  explicit Oracle(const char* path) {
    lib = dlopen(path, RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen %s: %s\n", path, dlerror()); exit(2); }
#define SYM(n) *(void**)(&n) = dlsym(lib, "lstmp_oracle_f32_" #n)
    SYM(create); SYM(destroy); SYM(set_params); SYM(get_params); SYM(get_grads); SYM(reset); SYM(propagate);
    SYM(backpropagate); SYM(update);
#undef SYM
  }
};

static double rel(const float* a, const float* b, size_t n) {
  double mx = 0, mb = 1e-30;
  for (size_t i = 0; i < n; i++) {
    mx = std::max(mx, (double)std::fabs(a[i] - b[i]));
    mb = std::max(mb, (double)std::fabs(b[i]));
  }
  return mx / mb;
}

int main(int argc, char** argv) {
  const char* oracle_path = argc > 1 ? argv[1] : "oracle/_build/liblstmp_oracle.so";
  Oracle O(oracle_path);
  const int I = 40, C = 64, R = 32, S = 4, T = 6;
  try {
    B200LstmProjectedStreams comp(I, R);
    std::istringstream cfg("<CellDim> 64 <NumStream> 4 <ParamScale> 0.1");
    comp.InitData(cfg);
    NnetTrainOptions opts;
    opts.learn_rate = 1e-3f;
    opts.momentum = 0.9f;
    comp.SetTrainOptions(opts);
    Vector<BaseFloat> flat;
    comp.GetParams(&flat);
    void* o = O.create(I, C, R, S);
    O.set_params(o, flat.Data());

    std::mt19937 rng(1234);
    std::normal_distribution<float> nd(0.f, 1.f);
    double worst = 0;
    for (int chunk = 0; chunk < 3; chunk++) {
      Matrix<BaseFloat> x(T * S, I), od(T * S, R), ref_out(T * S, R), ref_id(T * S, I);
      for (int i = 0; i < T * S; i++) {
        for (int j = 0; j < I; j++) x(i, j) = nd(rng);
        for (int j = 0; j < R; j++) od(i, j) = 0.1f * nd(rng);
      }
      std::vector<int> flags(S, 0);
      if (chunk == 2) flags[1] = flags[3] = 1;
      comp.Reset(flags);
      O.reset(o, flags.data(), S);
      CuMatrix<BaseFloat> cx(x), cod(od), cout_, cid;   // pitched device matrices (stride > cols)
      comp.Propagate(cx, &cout_);
      comp.Backpropagate(cx, cout_, cod, &cid);
      comp.Update(cx, cod);
      O.propagate(o, x.Data(), I, ref_out.Data(), R, T * S);
      O.backpropagate(o, x.Data(), I, od.Data(), R, ref_id.Data(), I, T * S, opts.momentum);
      O.update(o, opts.learn_rate);
      Matrix<BaseFloat> out, id;
      cout_.CopyToMat(&out);
      cid.CopyToMat(&id);
      Vector<BaseFloat> p, g;
      comp.GetParams(&p);
      comp.GetGradient(&g);
      std::vector<float> rp(p.Dim()), rg(p.Dim());
      O.get_params(o, rp.data());
      O.get_grads(o, rg.data());
      double e1 = rel(out.Data(), ref_out.Data(), (size_t)T * S * R), e2 = rel(id.Data(), ref_id.Data(), (size_t)T * S * I);
      double e3 = rel(g.Data(), rg.data(), rg.size()), e4 = rel(p.Data(), rp.data(), rp.size());
      printf("chunk %d: out %.2e in_diff %.2e corr %.2e params %.2e (stride in=%d out=%d)\n", chunk, e1, e2, e3, e4,
             cx.Stride(), cout_.Stride());
      worst = std::max(std::max(worst, e1), std::max(e2, std::max(e3, e4)));
    }
    // WriteData -> ReadData round trip in both Kaldi formats, and Copy()
    for (int binary = 0; binary <= 1; binary++) {
      std::stringstream ss;
      comp.WriteData(ss, binary != 0);
      B200LstmProjectedStreams twin(I, R);
      twin.ReadData(ss, binary != 0);
      Vector<BaseFloat> a, b;
      comp.GetParams(&a);
      twin.GetParams(&b);
      double e = rel(b.Data(), a.Data(), a.Dim());
      printf("write/read %s: %.2e\n", binary ? "binary" : "text", e);
      if (e > (binary ? 0.0 : 1e-5)) { printf("FAIL io\n"); return 1; }
    }
    Component* cp = comp.Copy();
    Vector<BaseFloat> a, b;
    comp.GetParams(&a);
    static_cast<B200LstmProjectedStreams*>(cp)->GetParams(&b);
    if (rel(b.Data(), a.Data(), a.Dim()) != 0.0) { printf("FAIL copy\n"); return 1; }
    delete cp;
    // error behaviour: rows not a multiple of NumStream asserts like LPS.h:225
    bool threw = false;
    try {
      CuMatrix<BaseFloat> bad(T * S + 1, I), o2;
      comp.Propagate(bad, &o2);
    } catch (const std::exception&) { threw = true; }
    if (!threw) { printf("FAIL assert\n"); return 1; }
    O.destroy(o);
    if (worst > 1e-4) { printf("FAIL parity %.3e\n", worst); return 1; }
    printf("PASS worst %.2e\n", worst);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 3;
  }
}
