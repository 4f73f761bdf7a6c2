// Test driver: kaldi/b200-affine-softmax-xent.h (Affine + Softmax + masked xent behind the C ABI) against a double
// precision host restatement of the [upstream] AffineTransform / Softmax semantics and of Xent::EvalMasked
// (google/nnet/nnet-loss.cc:76-164): posteriors, in_diff, parameters after two momentum updates, loss statistics.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "b200-affine-softmax-xent.h"

using namespace kaldi;
using namespace kaldi::nnet1;

int main() {
  try {
    const int I = 32, P = 200, F = 48;
    std::mt19937 rng(3);
    std::normal_distribution<float> nd(0.f, 1.f);
    Matrix<BaseFloat> W(P, I);
    Vector<BaseFloat> b(P);
    for (int r = 0; r < P; r++) {
      for (int c = 0; c < I; c++) W(r, c) = 0.1f * nd(rng);
      b(r) = -2.0f + 0.5f * nd(rng);
    }
    B200AffineSoftmaxXent tail(I, P, F);
    NnetTrainOptions opts;
    opts.learn_rate = 1e-2f;
    opts.momentum = 0.9f;
    tail.SetTrainOptions(opts);
    tail.SetParams(W, b);
    std::vector<double> Wd((size_t)P * I), bd(P), Wc((size_t)P * I, 0.0), bc(P, 0.0);
    for (int r = 0; r < P; r++) { for (int c = 0; c < I; c++) Wd[(size_t)r * I + c] = W(r, c); bd[r] = b(r); }
    double loss = 0, worst = 0;
    long frames = 0, correct = 0;
    for (int n = 0; n < 2; n++) {
      Matrix<BaseFloat> x(F, I);
      Vector<BaseFloat> mask(F);
      Posterior post(F);
      for (int t = 0; t < F; t++) {
        for (int c = 0; c < I; c++) x(t, c) = nd(rng);
        mask(t) = (t % 5 == 0) ? 0.f : 1.f;
        post[t].push_back(std::make_pair((int32)(rng() % P), 1.0f));
      }
      CuMatrix<BaseFloat> xd(x), yd, idd;
      tail.PropagateEval(xd, mask, post, &yd);
      tail.Backpropagate(xd, &idd);
      tail.Update();
      CU_SAFE_CALL(cudaDeviceSynchronize());
      Matrix<BaseFloat> y, ind;
      yd.CopyToMat(&y);
      idd.CopyToMat(&ind);
      // host restatement (double)
      std::vector<double> diff((size_t)F * P);
      for (int t = 0; t < F; t++) {
        std::vector<double> a(P);
        double mx = -1e300, sum = 0;
        for (int r = 0; r < P; r++) {
          double v = bd[r];
          for (int c = 0; c < I; c++) v += Wd[(size_t)r * I + c] * x(t, c);
          a[r] = v;
          mx = std::max(mx, v);
        }
        for (int r = 0; r < P; r++) { a[r] = std::exp(a[r] - mx); sum += a[r]; }
        int best = 0;
        for (int r = 0; r < P; r++) {
          a[r] /= sum;
          if (a[r] > a[best]) best = r;
          worst = std::max(worst, std::fabs(a[r] - y(t, r)));
          double tg = (post[t][0].first == r) ? 1.0 : 0.0;
          diff[(size_t)t * P + r] = (a[r] - tg) * mask(t);
        }
        if (mask(t) == 1.f) {
          loss -= std::log(a[post[t][0].first]);
          frames++;
          if (best == post[t][0].first) correct++;
        }
      }
      double imax = 0, ierr = 0;
      for (int t = 0; t < F; t++)
        for (int c = 0; c < I; c++) {
          double v = 0;
          for (int r = 0; r < P; r++) v += diff[(size_t)t * P + r] * Wd[(size_t)r * I + c];
          imax = std::max(imax, std::fabs(v));
          ierr = std::max(ierr, std::fabs(v - ind(t, c)));
        }
      if (ierr > 1e-4 * imax) { printf("FAIL: in_diff error %g of %g\n", ierr, imax); return 1; }
      for (int r = 0; r < P; r++) {
        double gb = 0;
        for (int t = 0; t < F; t++) gb += diff[(size_t)t * P + r];
        bc[r] = gb + 0.9 * bc[r];
        for (int c = 0; c < I; c++) {
          double g = 0;
          for (int t = 0; t < F; t++) g += diff[(size_t)t * P + r] * x(t, c);
          Wc[(size_t)r * I + c] = g + 0.9 * Wc[(size_t)r * I + c];
        }
      }
      for (int r = 0; r < P; r++) {
        bd[r] -= 1e-2 * bc[r];
        for (int c = 0; c < I; c++) Wd[(size_t)r * I + c] -= 1e-2 * Wc[(size_t)r * I + c];
      }
    }
    if (worst > 1e-5) { printf("FAIL: posterior error %g\n", worst); return 1; }
    Matrix<BaseFloat> W2;
    Vector<BaseFloat> b2;
    tail.GetParams(&W2, &b2);
    double perr = 0, pmax = 0;
    for (int r = 0; r < P; r++) {
      perr = std::max(perr, std::fabs(b2(r) - bd[r]));
      pmax = std::max(pmax, std::fabs(bd[r]));
      for (int c = 0; c < I; c++) perr = std::max(perr, std::fabs(W2(r, c) - Wd[(size_t)r * I + c]));
    }
    if (perr > 1e-4 * pmax) { printf("FAIL: parameter error %g of %g\n", perr, pmax); return 1; }
    lstmp_b200_xent_stats_t st = tail.Stats();
    if (st.frames != frames || st.correct != correct || std::fabs(st.loss - loss) > 1e-4 * std::fabs(loss)) {
      printf("FAIL: stats frames %lld/%ld correct %lld/%ld loss %g/%g\n", (long long)st.frames, frames, (long long)st.correct,
             correct, st.loss, loss);
      return 1;
    }
    printf("%s\nPASS\n", tail.Report().c_str());
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "tail_test: %s\n", e.what());
    return 2;
  }
}
