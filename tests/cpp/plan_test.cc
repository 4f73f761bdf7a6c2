// CPU test of the tile-image GEMM's host-side planning arithmetic (kaldi-lstm_b200/csrc/lstmp_gemm_plan.h): work-item
// tiling and the split-K plan of a group launch.  No CUDA, no GPU: g++ only.
//   (1) properties of the tiling and of the chosen plans (bounds, workspace limit, optimal under the cost model);
//   (2) the plans of the shapes the engine launches (cfg3 both layers, cfg4 per GPU, cfg5, the tail), printed with the
//       load of the most loaded CTA next to the average: items are dealt round-robin (item w runs on CTA w mod 148), so
//       a group of few, long items stays imbalanced whatever the split factors (cfg3 layer 2: 72 vs 46 K blocks -- the
//       measured 59 us of that launch are 0.82 us per K block on its critical CTAs, profiles/r2_ncu_digest.md);
//   (3) an independent brute-force search over the same candidates, written against the definitions rather than the
//       header's helpers, agrees on 2000 random groups.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#include "../../kaldi-lstm_b200/csrc/lstmp_gemm_plan.h"

using namespace lstmp::hlplan;

static int g_fail = 0;
#define CHECK(cond, ...)                      \
  do {                                        \
    if (!(cond)) {                            \
      ++g_fail;                               \
      printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);                    \
      printf("\n");                           \
    }                                         \
  } while (0)

// ---- independent restatement: enumerate every candidate, simulate the round-robin deal item by item ----------------
struct Item {
  int len;
};
static double brute_cost(const Shape* d, const int* want, int n, int nsm, size_t* wsf_out) {
  struct Prod {
    int nkt, tiles, splits, per;
    size_t mn;
  };
  std::vector<Prod> P;
  size_t wsf = 0;
  double traffic = 0;
  for (int i = 0; i < n; ++i) {
    Prod p;
    p.nkt = (d[i].K + 63) / 64;
    p.tiles = ((d[i].M + 127) / 128) * ((d[i].N + 127) / 128);
    p.per = p.nkt;
    p.splits = 1;
    if (want[i] > 1) {
      p.per = (p.nkt + want[i] - 1) / want[i];
      p.splits = (p.nkt + p.per - 1) / p.per;
    }
    p.mn = (size_t)d[i].M * d[i].N;
    if (p.splits > 1) {
      wsf += p.splits * p.mn;
      traffic += (double)(p.splits + 1) * p.mn * 4.0;
    }
    P.push_back(p);
  }
  *wsf_out = wsf;
  std::stable_sort(P.begin(), P.end(), [](const Prod& a, const Prod& b) { return a.per > b.per; });
  std::vector<Item> items;
  for (const Prod& p : P)
    for (int z = 0; z < p.splits; ++z)
      for (int t = 0; t < p.tiles; ++t) items.push_back(Item{std::min(p.per, p.nkt - z * p.per) + 1});
  std::vector<int> load((size_t)nsm, 0);
  for (size_t w = 0; w < items.size(); ++w) load[w % (size_t)nsm] += items[w].len;
  const int ms = *std::max_element(load.begin(), load.end());
  return 0.4 * ms + (traffic > 0 ? 3.0 + traffic / 3.0e6 : 0.0);
}
static void brute_plan(const Shape* d, int n, int nsm, bool have_ws, size_t ws_floats, int* out) {
  int smax[kMaxGroup], cur[kMaxGroup];
  for (int i = 0; i < n; ++i) {
    const int nkt = (d[i].K + 63) / 64;
    smax[i] = (have_ws && d[i].N % 4 == 0) ? std::max(1, std::min(6, nkt / 2)) : 1;
    cur[i] = out[i] = 1;
  }
  double best = 1e30;
  for (;;) {
    size_t wsf = 0;
    const double c = brute_cost(d, cur, n, nsm, &wsf);
    if (wsf <= ws_floats && c < best) {
      best = c;
      for (int i = 0; i < n; ++i) out[i] = cur[i];
    }
    int i = 0;
    while (i < n && ++cur[i] > smax[i]) cur[i++] = 1;
    if (i == n) break;
  }
}

static void show(const char* name, const Shape* d, int n, int nsm, size_t ws_floats) {
  int sp[kMaxGroup];
  plan_group(d, n, nsm, true, ws_floats, sp);
  Tiling P[kMaxGroup];
  long long total = 0;
  int items = 0;
  for (int i = 0; i < n; ++i) {
    P[i] = tiling(d[i].M, d[i].N, d[i].K, sp[i]);
    CHECK(P[i].splits >= 1 && P[i].splits <= 6, "%s: splits %d", name, P[i].splits);
    for (int z = 0; z < P[i].splits; ++z)
      total += (long long)P[i].ntm * P[i].ntn * (std::min(P[i].kt_per_split, P[i].nkt - z * P[i].kt_per_split) + 1);
    items += P[i].ntm * P[i].ntn * P[i].splits;
  }
  std::stable_sort(P, P + n, [](const Tiling& x, const Tiling& y) { return x.kt_per_split > y.kt_per_split; });
  std::vector<int> load;
  const int ms = makespan(P, n, nsm, load);
  const double avg = (double)total / nsm;
  printf("%-28s splits", name);
  for (int i = 0; i < n; ++i) printf(" %d", sp[i]);
  printf("  items %4d  most loaded CTA %3d K blocks, average %.1f (%.2f)\n", items, ms, avg, ms / avg);
  CHECK(ms >= avg - 1e-9, "%s: makespan %d below the average %.1f", name, ms, avg);
  // the plan is never worse (under the model) than the same group without split-K
  int ones[kMaxGroup] = {1, 1, 1, 1};
  size_t w0 = 0, w1 = 0;
  CHECK(group_cost(d, sp, n, nsm, &w0, load) <= group_cost(d, ones, n, nsm, &w1, load) + 1e-9, "%s: plan worse than none", name);
  CHECK(w0 <= ws_floats, "%s: %zu floats of workspace", name, w0);
}

int main() {
  // ---- (1) tiling
  for (int K : {1, 40, 64, 65, 512, 1280, 3200, 16624})
    for (int want = 1; want <= 8; ++want) {
      const Tiling t = tiling(1280, 3200, K, want);
      CHECK(t.nkt == (K + 63) / 64 && t.ntm == 10 && t.ntn == 25, "tile counts K=%d", K);
      CHECK(t.splits >= 1 && t.splits <= std::max(1, want), "splits %d for want %d", t.splits, want);
      CHECK(t.kt_per_split * t.splits >= t.nkt, "K=%d want=%d: slices do not cover K", K, want);
      CHECK((t.splits - 1) * t.kt_per_split < t.nkt, "K=%d want=%d: an empty slice", K, want);
    }
  {
    const Tiling t = tiling(1, 1, 1, 1);
    CHECK(t.nkt == 1 && t.ntm == 1 && t.ntn == 1 && t.splits == 1 && t.kt_per_split == 1, "smallest product");
  }
  // ---- (2) the engine's groups (lstmp_engine.cu: in_diff, G(w_gifo_x), G(w_gifo_r), G(w_r_m)), 148 SMs, 4 Mi floats
  const size_t WS = (size_t)4 << 20;
  const Shape l2[4] = {{1280, 512, 3200}, {3200, 512, 1280}, {3200, 512, 1280}, {512, 800, 1280}};
  const Shape l1[3] = {{3200, 40, 1280}, {3200, 512, 1280}, {512, 800, 1280}};
  const Shape c4[3] = {{3200, 40, 640}, {3200, 512, 640}, {512, 800, 640}};
  const Shape c5[3] = {{8192, 40, 1280}, {8192, 1024, 1280}, {1024, 2048, 1280}};
  const Shape deep[1] = {{640, 512, 16624}};
  show("cfg3 layer 2 (4 products)", l2, 4, 148, WS);
  show("cfg3 layer 1 (3 products)", l1, 3, 148, WS);
  show("cfg4 per GPU, S=32", c4, 3, 148, WS);
  show("cfg5 per GPU", c5, 3, 148, WS);
  show("tail in_diff, K=16624", deep, 1, 148, WS);
  {
    int sp[4];
    plan_group(l2, 4, 148, false, WS, sp);   // no workspace: no split-K at all
    CHECK(sp[0] == 1 && sp[1] == 1 && sp[2] == 1 && sp[3] == 1, "split-K without a workspace");
    plan_group(l2, 4, 148, true, 0, sp);     // workspace of zero floats: likewise
    CHECK(sp[0] == 1 && sp[1] == 1 && sp[2] == 1 && sp[3] == 1, "split-K with an empty workspace");
    const Shape odd[2] = {{640, 514, 16624}, {640, 512, 16624}};   // N % 4 != 0: that product is never split
    plan_group(odd, 2, 148, true, WS, sp);
    CHECK(sp[0] == 1, "N %% 4 != 0 was split (%d)", sp[0]);
    plan_group(deep, 1, 148, true, WS, sp);   // 20 output tiles, 260 K blocks each: split as far as allowed
    CHECK(tiling(deep[0].M, deep[0].N, deep[0].K, sp[0]).splits == 6, "a 20-tile, 260-block product got %d slices", sp[0]);
  }
  // ---- (3) random groups: the header's plan == an independent brute force; fits the workspace; never worse than no split
  std::mt19937 rng(1234);
  auto rnd = [&](int lo, int hi) { return lo + (int)(rng() % (unsigned)(hi - lo + 1)); };
  for (int it = 0; it < 2000; ++it) {
    const int n = rnd(1, 4), nsm = (it % 5 == 0) ? rnd(1, 200) : 148;
    Shape d[kMaxGroup];
    for (int i = 0; i < n; ++i) d[i] = Shape{rnd(1, 4000), rnd(1, 1200), rnd(1, 6000)};
    const size_t ws = (it % 7 == 0) ? (size_t)rnd(0, 1 << 20) : WS;
    const bool have = it % 11 != 0;
    int a[kMaxGroup], b[kMaxGroup];
    plan_group(d, n, nsm, have, ws, a);
    brute_plan(d, n, nsm, have, ws, b);
    for (int i = 0; i < n; ++i) CHECK(a[i] == b[i], "group %d product %d: plan %d vs brute force %d", it, i, a[i], b[i]);
    size_t wsf = 0, wsf1 = 0;
    std::vector<int> load;
    const double c = group_cost(d, a, n, nsm, &wsf, load);
    int ones[kMaxGroup] = {1, 1, 1, 1};
    const double c1 = group_cost(d, ones, n, nsm, &wsf1, load);
    CHECK(wsf <= ws, "group %d: plan needs %zu floats of %zu", it, wsf, ws);
    CHECK(c <= c1 + 1e-9, "group %d: plan costs %.2f, no split-K %.2f", it, c, c1);
    CHECK(wsf1 == 0, "no split-K needs no workspace");
  }
  printf(g_fail ? "FAILED (%d)\n" : "PASS\n", g_fail);
  return g_fail ? 1 : 0;
}
