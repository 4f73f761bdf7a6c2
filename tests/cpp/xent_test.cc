// C++ test of the Kaldi-side mirror of Xent::EvalMasked (kaldi-lstm_b200/kaldi/b200-nnet-loss.h) against a dense
// host restatement of google/nnet/nnet-loss.cc:76-164 written here (tests only).  Needs a B200.  Prints "PASS".
#include <cmath>
#include <cstdio>
#include <random>

#include "b200-nnet-loss.h"

using namespace kaldi;
using namespace kaldi::nnet1;

int main() {
  try {
    const int rows = 48, P = 301;
    std::mt19937 rng(7);
    std::normal_distribution<float> nd(0.f, 2.f);
    Matrix<BaseFloat> y(rows, P);
    Posterior post(rows);
    Vector<BaseFloat> mask(rows);
    for (int t = 0; t < rows; t++) {
      double z = 0;
      std::vector<float> e(P);
      for (int c = 0; c < P; c++) { e[c] = std::exp(nd(rng)); z += e[c]; }
      for (int c = 0; c < P; c++) y(t, c) = (float)(e[c] / z);
      mask(t) = (t % 4 == 3) ? 0.f : 1.f;
      if (t % 9 == 8) continue;  // empty posterior
      int p = rng() % P;
      post[t].push_back(std::make_pair(p, 0.75f));
      post[t].push_back(std::make_pair(t % 5 == 0 ? p : (int)(rng() % P), 0.25f));  // sometimes a duplicate pdf
    }
    // dense host reference
    std::vector<float> tgt((size_t)rows * P, 0.f), ref((size_t)rows * P);
    double loss = 0, ent = 0;
    long long correct = 0, frames = 0;
    for (int t = 0; t < rows; t++) {
      for (auto& pw : post[t]) tgt[(size_t)t * P + pw.first] += pw.second;
      int ay = -1, at = -1;
      float my = -1e21f, mt = -1e21f;
      for (int c = 0; c < P; c++) {
        float yy = y(t, c), tt = tgt[(size_t)t * P + c];
        ref[(size_t)t * P + c] = (yy - tt) * mask(t);
        if (my < yy) { my = yy; ay = c; }
        if (mt < tt) { mt = tt; at = c; }
        loss -= (double)((std::log(yy) * tt) * mask(t));
        ent -= (double)((std::log(tt + 1e-20f) * tt) * mask(t));
      }
      if (mask(t) == 1.f && ay == at) correct++;
      frames += (long long)mask(t);
    }
    CuMatrix<BaseFloat> net_out(y), diff;
    B200Xent xent;
    xent.EvalMasked(mask, net_out, post, &diff);
    xent.EvalMasked(mask, net_out, post, &diff);  // statistics accumulate over calls
    Matrix<BaseFloat> got;
    diff.CopyToMat(&got);
    for (int t = 0; t < rows; t++)
      for (int c = 0; c < P; c++)
        if (got(t, c) != ref[(size_t)t * P + c]) { printf("FAIL diff (%d,%d): %g vs %g\n", t, c, got(t, c), ref[(size_t)t * P + c]); return 1; }
    B200Xent::Stats s = xent.GetStats();
    if (std::fabs(s.loss - 2 * loss) > 1e-5 * std::fabs(loss) || std::fabs(s.entropy - 2 * ent) > 1e-5 * std::fabs(ent) + 1e-9 ||
        s.correct != 2 * correct || s.frames != 2 * frames) {
      printf("FAIL stats: %.9g/%.9g %.9g/%.9g %lld/%lld %lld/%lld\n", s.loss, 2 * loss, s.entropy, 2 * ent, s.correct,
             2 * correct, s.frames, 2 * frames);
      return 1;
    }
    // a bigger batch re-creates the engine and keeps the statistics
    Matrix<BaseFloat> y2(2 * rows, P);
    Posterior post2(2 * rows);
    Vector<BaseFloat> mask2(2 * rows);
    for (int t = 0; t < 2 * rows; t++) {
      for (int c = 0; c < P; c++) y2(t, c) = y(t % rows, c);
      post2[t] = post[t % rows];
      mask2(t) = mask(t % rows);
    }
    CuMatrix<BaseFloat> net2(y2);
    xent.EvalMasked(mask2, net2, post2, &diff);
    s = xent.GetStats();
    if (s.frames != 4 * frames || s.correct != 4 * correct) { printf("FAIL regrow\n"); return 1; }
    // pdf-id outside the network output: KALDI_ERR (nnet-loss.cc:88-91)
    bool threw = false;
    try {
      Posterior bad(rows);
      bad[0].push_back(std::make_pair(P, 1.0f));
      xent.EvalMasked(mask, net_out, bad, &diff);
    } catch (const std::exception&) { threw = true; }
    if (!threw) { printf("FAIL pdf range\n"); return 1; }
    printf("%s\nPASS\n", xent.Report().c_str());
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 3;
  }
}
