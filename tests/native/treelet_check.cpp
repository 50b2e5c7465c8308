// Host-side check of the builder's binary-tree stages (ospray_b200/csrc/cuda/treelet.h compiles for the host too):
// Morton sort -> Karras radix tree -> refit -> treelet restructuring passes run in the same bottom-up order as
// treeletKernel (one "thread" per leaf climbing with arrival counters, executed one after the other = one legal
// interleaving of the GPU schedule). Validates the tree after every pass and prints the SAH cost.
//   usage: treelet_check <n> <distribution 0..2> <passes>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../ospray_b200/csrc/cuda/treelet.h"
using namespace oc;

static uint64_t expand21(uint64_t v)
{
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

struct Tree
{
  int n;
  std::vector<BinNode> nodes;
  std::vector<uint32_t> parent;
  std::vector<Aabb> boxes;
};

static double sahCost(const Tree &t)
{
  double c = 0;
  for (int i = 0; i < t.n - 1; i++) c += halfArea(t.boxes[i]);
  return c / halfArea(t.boxes[0]);
}

// returns 0 when the tree is a valid binary tree over the n leaves with exact boxes and leaf counts
static int validate(const Tree &t)
{
  const uint32_t leafBase = t.n - 1;
  std::vector<int> seen(t.n, 0), innerSeen(t.n - 1, 0);
  std::vector<uint32_t> order, stack{0};
  if (t.parent[0] != 0xFFFFFFFFu) return 1;
  while (!stack.empty()) {
    const uint32_t id = stack.back(); stack.pop_back();
    if (id >= leafBase) { if (id - leafBase >= (uint32_t)t.n || seen[id - leafBase]++) return 2; continue; }
    if (innerSeen[id]++) return 3;
    order.push_back(id);
    const BinNode &b = t.nodes[id];
    if (t.parent[b.left] != id || t.parent[b.right] != id) return 4;
    stack.push_back(b.left); stack.push_back(b.right);
  }
  for (int j = 0; j < t.n; j++) if (!seen[j]) return 5;
  for (int i = 0; i < t.n - 1; i++) if (!innerSeen[i]) return 6;
  std::vector<uint32_t> cnt(2 * t.n - 1, 1);
  for (size_t k = order.size(); k-- > 0;) { // children before parents
    const uint32_t id = order[k];
    const BinNode &b = t.nodes[id];
    cnt[id] = cnt[b.left] + cnt[b.right];
    if (b.last - b.first + 1 != cnt[id]) return 7;
    const Aabb u = unionOf(t.boxes[b.left], t.boxes[b.right]);
    if (memcmp(&u, &t.boxes[id], sizeof(Aabb)) != 0) return 8;
  }
  return 0;
}

static int treeletPass(Tree &t, int gamma)
{
  std::vector<uint32_t> flags(t.n - 1, 0);
  int rewritten = 0;
  for (int j = 0; j < t.n; j++) {
    uint32_t p = t.parent[t.n - 1 + j];
    while (p != 0xFFFFFFFFu) {
      if (flags[p]++ == 0) break;
      const BinNode b = t.nodes[p];
      if (b.last - b.first + 1 >= (uint32_t)gamma) rewritten += optimizeTreelet(p, t.n, t.nodes.data(), t.parent.data(), t.boxes.data());
      p = t.parent[p];
    }
  }
  return rewritten;
}

// The collapse of bvh_build.cu (collapseKernel) on the host: children with the largest surface area are opened first until
// the node has 4 slots; subtrees of <= leafMax leaves become leaves. Returns, through `need`, the deepest traversal stack a
// depth-first walk can build = max over root-to-leaf paths of the sum of (children - 1) of the wide nodes on the path
// (every level pushes all but the child it descends into), and the wide depth.
static void collapseStats(const Tree &t, uint32_t bin, uint32_t leafMax, int depth, int pushed, int &maxDepth, int &need)
{
  const uint32_t leafBase = t.n - 1;
  auto size = [&](uint32_t id) { return id >= leafBase ? 1u : t.nodes[id].last - t.nodes[id].first + 1u; };
  if (size(bin) <= leafMax) { maxDepth = std::max(maxDepth, depth); need = std::max(need, pushed); return; }
  uint32_t slot[4];
  int cnt = 0;
  slot[cnt++] = t.nodes[bin].left; slot[cnt++] = t.nodes[bin].right;
  while (cnt < 4) {
    int best = -1; float bestA = -1.f;
    for (int k = 0; k < cnt; k++) {
      if (size(slot[k]) <= leafMax) continue;
      const float a = halfArea(t.boxes[slot[k]]);
      if (a > bestA) { bestA = a; best = k; }
    }
    if (best < 0) break;
    const uint32_t open = slot[best];
    slot[best] = t.nodes[open].left; slot[cnt++] = t.nodes[open].right;
  }
  for (int k = 0; k < cnt; k++) collapseStats(t, slot[k], leafMax, depth + 1, pushed + cnt - 1, maxDepth, need);
}

int main(int argc, char **argv)
{
  const int n = argc > 1 ? atoi(argv[1]) : 20000, dist = argc > 2 ? atoi(argv[2]) : 0, passes = argc > 3 ? atoi(argv[3]) : 3;
  std::mt19937 rng(1234 + dist);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  std::vector<Aabb> prim(n);
  for (int i = 0; i < n; i++) {
    float c[3], e;
    if (dist == 0) { for (float &x : c) x = U(rng); e = 0.01f * U(rng); }                                   // small boxes, uniform
    else if (dist == 1) { for (float &x : c) x = U(rng); e = (i % 50 == 0) ? 0.4f * U(rng) : 0.004f * U(rng); } // a few huge ones
    else if (dist == 3) { // jittered lattice of boxes, 12 triangles each (the C3 mesh of tools/bench_workloads.py in small)
      const int b = i / 12, side = std::max(1, (int)cbrtf((float)(n / 12)));
      const float cell = 1.f / side;
      const float bc[3] = {(b % side + 0.5f) * cell, ((b / side) % side + 0.5f) * cell, (b / (side * side) + 0.5f) * cell};
      const int face = (i % 12) / 2;
      for (int k = 0; k < 3; k++) c[k] = bc[k] + (k == face / 2 ? ((face & 1) ? 0.3f : -0.3f) * cell : 0.1f * cell * (U(rng) - 0.5f));
      e = 0.3f * cell;
      for (int k = 0; k < 3; k++) { const float ek = (k == face / 2) ? 1e-4f * cell : e * (0.5f + 0.5f * U(rng)); prim[i].lo[k] = c[k] - ek; prim[i].hi[k] = c[k] + ek; }
      continue;
    }
    else { const int cl = i % 16; for (int k = 0; k < 3; k++) c[k] = 0.0625f * ((cl * (k + 3)) % 16) + 0.02f * U(rng); e = 0.002f; if (i % 7 == 0) c[1] = c[0]; } // clusters + duplicates
    if (dist == 2 && i % 5 == 0 && i) { prim[i] = prim[i - 1]; continue; }                                        // identical keys
    for (int k = 0; k < 3; k++) { prim[i].lo[k] = c[k] - e * U(rng); prim[i].hi[k] = c[k] + e * U(rng); }
  }
  Aabb sb = prim[0];
  for (const Aabb &b : prim) sb = unionOf(sb, b);
  std::vector<std::pair<uint64_t, uint32_t>> ks(n);
  for (int i = 0; i < n; i++) {
    uint64_t key = 0;
    for (int a = 0; a < 3; a++) {
      const float ext = sb.hi[a] - sb.lo[a], c = 0.5f * (prim[i].lo[a] + prim[i].hi[a]);
      float f = ext > 0.f ? (c - sb.lo[a]) / ext : 0.f;
      f = fminf(fmaxf(f * 2097152.f, 0.f), 2097151.f);
      key |= expand21((uint64_t)f) << (2 - a);
    }
    ks[i] = {key, (uint32_t)i};
  }
  std::stable_sort(ks.begin(), ks.end(), [](auto &a, auto &b) { return a.first < b.first; });
  std::vector<uint64_t> keys(n);
  for (int i = 0; i < n; i++) keys[i] = ks[i].first;

  Tree t;
  t.n = n;
  t.nodes.resize(n - 1); t.parent.assign(2 * n - 1, 0xFFFFFFFFu); t.boxes.resize(2 * n - 1);
  for (int i = 0; i < n - 1; i++) {
    t.nodes[i] = karrasNode(keys.data(), n, i);
    t.parent[t.nodes[i].left] = i; t.parent[t.nodes[i].right] = i;
  }
  t.parent[0] = 0xFFFFFFFFu;
  { // refit bottom-up with arrival counters, as refitKernel does
    std::vector<uint32_t> flags(n - 1, 0);
    for (int j = 0; j < n; j++) {
      t.boxes[n - 1 + j] = prim[ks[j].second];
      uint32_t p = t.parent[n - 1 + j];
      while (p != 0xFFFFFFFFu) {
        if (flags[p]++ == 0) break;
        t.boxes[p] = unionOf(t.boxes[t.nodes[p].left], t.boxes[t.nodes[p].right]);
        p = t.parent[p];
      }
    }
  }
  int rc = validate(t);
  printf("{\"n\": %d, \"dist\": %d, \"valid\": [%d", n, dist, rc);
  std::vector<double> cost{sahCost(t)};
  std::vector<int> rew;
  for (int p = 0; p < passes; p++) {
    rew.push_back(treeletPass(t, 7));
    const int v = validate(t);
    rc |= v;
    printf(", %d", v);
    cost.push_back(sahCost(t));
  }
  int wideDepth = 0, stackNeed = 0;
  if (n > 1) collapseStats(t, 0, 2, 0, 0, wideDepth, stackNeed);
  printf("], \"wide_depth\": %d, \"stack_need\": %d, \"cost\": [", wideDepth, stackNeed);
  for (size_t k = 0; k < cost.size(); k++) printf("%s%.4f", k ? ", " : "", cost[k]);
  printf("], \"rewritten\": [");
  for (size_t k = 0; k < rew.size(); k++) printf("%s%d", k ? ", " : "", rew[k]);
  printf("]}\n");
  return rc;
}
