// Host-only planning arithmetic of the tile-image GEMM (lstmp_gemm_hl.cu): how a product is cut into work items and
// which split-K factors a GROUP of products gets.  No CUDA in here: tests/cpp/plan_test.cc checks it on the CPU.
#pragma once
#include <stddef.h>

#include <algorithm>
#include <vector>

namespace lstmp {
namespace hlplan {
constexpr int kBM = 128, kBN = 128, kBK = 64;  // output tile and K block of gemm_hl_kernel
constexpr int kMaxGroup = 4;                   // products per group launch

// Work items of one product C[M x N] over K: ntm x ntn output tiles x `splits` K slices of kt_per_split K blocks.
// `want_splits` is a wish: the slices are evened out and empty ones dropped.
struct Tiling {
  int nkt, ntm, ntn, splits, kt_per_split;
};
inline Tiling tiling(int M, int N, int K, int want_splits) {
  Tiling t;
  t.nkt = (K + kBK - 1) / kBK;
  t.ntm = (M + kBM - 1) / kBM;
  t.ntn = (N + kBN - 1) / kBN;
  t.kt_per_split = t.nkt;
  t.splits = 1;
  if (want_splits > 1) {
    t.kt_per_split = (t.nkt + want_splits - 1) / want_splits;
    t.splits = (t.nkt + t.kt_per_split - 1) / t.kt_per_split;
  }
  return t;
}

// K blocks on the most loaded CTA when the items of the products (in the given order) are dealt round-robin to
// `nsm` CTAs (item w of the launch runs on CTA w mod nsm, as the persistent kernel walks them).
inline int makespan(const Tiling* P, int n, int nsm, std::vector<int>& load) {
  load.assign((size_t)nsm, 0);
  int w = 0;
  for (int i = 0; i < n; ++i)
    for (int z = 0; z < P[i].splits; ++z) {
      const int len = std::min(P[i].kt_per_split, P[i].nkt - z * P[i].kt_per_split) + 1;   // (+1: per-item overhead)
      for (int t = 0; t < P[i].ntm * P[i].ntn; ++t, ++w) load[(size_t)(w % nsm)] += len;
    }
  return *std::max_element(load.begin(), load.end());
}

struct Shape {
  int M, N, K;
};
// Cost model of one candidate: 0.4 us per K block on the most loaded CTA (64 KB from L2 per block), products ordered
// longest items first; split-K adds one reduce launch (3 us) and its traffic at 3 TB/s.
inline double group_cost(const Shape* d, const int* want, int n, int nsm, size_t* ws_floats_needed, std::vector<int>& load) {
  Tiling P[kMaxGroup] = {};
  size_t wsf = 0;
  double traffic = 0;
  for (int i = 0; i < n; ++i) {
    P[i] = tiling(d[i].M, d[i].N, d[i].K, want[i]);
    if (P[i].splits > 1) {
      wsf += (size_t)P[i].splits * d[i].M * d[i].N;
      traffic += (double)(P[i].splits + 1) * d[i].M * d[i].N * 4.0;
    }
  }
  *ws_floats_needed = wsf;
  std::stable_sort(P, P + n, [](const Tiling& x, const Tiling& y) { return x.kt_per_split > y.kt_per_split; });
  return 0.4 * makespan(P, n, nsm, load) + (traffic > 0 ? 3.0 + traffic / 3.0e6 : 0.0);
}

// Split-K wishes (1..6, at most nkt / 2; 1 when there is no workspace or N is not a multiple of 4: the reduce is
// vectorised) of the n <= kMaxGroup products of a group: the cheapest candidate whose partial sums fit in `ws_floats`.
// Exhaustive (<= 6^4 candidates); the first of equally cheap candidates wins, product 0's factor varying fastest.
inline void plan_group(const Shape* d, int n, int nsm, bool have_ws, size_t ws_floats, int* splits) {
  int smax[kMaxGroup], cur[kMaxGroup];
  for (int i = 0; i < n; ++i) {
    const int nkt = (d[i].K + kBK - 1) / kBK;
    smax[i] = (have_ws && (d[i].N & 3) == 0) ? std::max(1, std::min(6, nkt / 2)) : 1;
    cur[i] = splits[i] = 1;
  }
  double best_cost = 1e30;
  std::vector<int> load;
  for (;;) {
    size_t wsf = 0;
    const double cost = group_cost(d, cur, n, nsm, &wsf, load);
    if (wsf <= ws_floats && cost < best_cost) {
      best_cost = cost;
      for (int i = 0; i < n; ++i) splits[i] = cur[i];
    }
    int i = 0;
    while (i < n && ++cur[i] > smax[i]) cur[i++] = 1;
    if (i == n) break;
  }
}
}  // namespace hlplan
}  // namespace lstmp
