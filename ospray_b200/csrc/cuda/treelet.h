// Binary-BVH pieces shared by the CUDA builder (bvh_build.cu) and its host-side check (tests/treelet_check.cpp):
// the Karras 2012 radix-tree node and the treelet restructuring of Karras & Aila 2013 ("Fast parallel construction of
// high-quality bounding volume hierarchies", HPG 2013): every subtree root with >= 7 leaves below it grows a treelet
// of 7 leaves (largest surface area opened first), solves the optimal binary topology over those 7 leaves by dynamic
// programming over the 127 leaf subsets with the surface-area heuristic as the cost, and rewrites the 6 internal nodes
// in place. This is the SAH refinement of the Morton-code LBVH that stands in for Embree's SAH builder behind
// rtcCommitScene (modules/cpu/common/Group.cpp:129-142, World.cpp:167-171).
//
// Everything here is plain C++ that compiles for the host as well (no CUDA-only types): `OC_LD*` read through L2 on the
// device because neighbouring threads of the bottom-up pass wrote the data.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define OC_HD __host__ __device__ __forceinline__
#else
#define OC_HD inline
#endif

namespace oc {

struct Aabb { float lo[3], hi[3]; };

// node ids: internal i in [0, n-1), leaf j as (n-1)+j. first/last: after the Karras pass the range of sorted leaves
// below the node; after a treelet rewrite only last-first+1 (= number of leaves below) is meaningful.
struct BinNode { uint32_t left, right, first, last; };

OC_HD int ocClz64(uint64_t v)
{
#ifdef __CUDA_ARCH__
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
OC_HD int ocClz32(uint32_t v)
{
#ifdef __CUDA_ARCH__
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}

OC_HD int deltaKey(const uint64_t *keys, int n, int i, int j)
{
  if (j < 0 || j >= n) return -1;
  const uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + ocClz32((uint32_t)i ^ (uint32_t)j);
  return ocClz64(a ^ b);
}

// internal node i of the radix tree over n sorted keys (Karras 2012, section 4)
OC_HD BinNode karrasNode(const uint64_t *keys, int n, int i)
{
  const int d = (deltaKey(keys, n, i, i + 1) - deltaKey(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = deltaKey(keys, n, i, i - d);
  int lmax = 2;
  while (deltaKey(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (deltaKey(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = deltaKey(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (deltaKey(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + (d < 0 ? d : 0);
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  BinNode b;
  b.left = (lo == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
  b.right = (hi == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
  b.first = (uint32_t)lo;
  b.last = (uint32_t)hi;
  return b;
}

OC_HD float halfArea(const Aabb &b)
{
  const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return dx * dy + dy * dz + dz * dx;
}
OC_HD Aabb unionOf(const Aabb &a, const Aabb &b)
{
  Aabb u;
  for (int k = 0; k < 3; k++) {
    u.lo[k] = fminf(a.lo[k], b.lo[k]);
    u.hi[k] = fmaxf(a.hi[k], b.hi[k]);
  }
  return u;
}

OC_HD Aabb ocLdBox(const Aabb *p)
{
  Aabb b;
#ifdef __CUDA_ARCH__
  const float *f = reinterpret_cast<const float *>(p);
  for (int k = 0; k < 3; k++) { b.lo[k] = __ldcg(f + k); b.hi[k] = __ldcg(f + 3 + k); }
#else
  b = *p;
#endif
  return b;
}
OC_HD BinNode ocLdNode(const BinNode *p)
{
#ifdef __CUDA_ARCH__
  const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p));
  BinNode b; b.left = v.x; b.right = v.y; b.first = v.z; b.last = v.w;
  return b;
#else
  return *p;
#endif
}

#define OC_TREELET_LEAVES 7
#define OC_TREELET_SUBSETS (1 << OC_TREELET_LEAVES)

// Restructures the treelet rooted at internal node `root` (>= 7 leaves below it). Touches only nodes inside root's
// subtree; root keeps its id, box, parent and leaf count. Returns true when the topology was rewritten.
OC_HD bool optimizeTreelet(uint32_t root, int n, BinNode *nodes, uint32_t *parent, Aabb *boxes)
{
  const uint32_t leafBase = (uint32_t)(n - 1);
  const int NL = OC_TREELET_LEAVES;
  uint32_t leaf[OC_TREELET_LEAVES], inner[OC_TREELET_LEAVES - 1];
  Aabb lb[OC_TREELET_LEAVES];
  float la[OC_TREELET_LEAVES];
  const BinNode r = ocLdNode(nodes + root);
  int nl = 2, ni = 1;
  inner[0] = root;
  leaf[0] = r.left; leaf[1] = r.right;
  lb[0] = ocLdBox(boxes + r.left); lb[1] = ocLdBox(boxes + r.right);
  la[0] = halfArea(lb[0]); la[1] = halfArea(lb[1]);
  float curCost = 0.f; // surface area of the internal nodes below the root as they are now
  while (nl < NL) {
    int best = -1;
    float bestA = -1.f;
    for (int k = 0; k < nl; k++)
      if (leaf[k] < leafBase && la[k] > bestA) { bestA = la[k]; best = k; }
    if (best < 0) return false;
    const uint32_t id = leaf[best];
    const BinNode b = ocLdNode(nodes + id);
    inner[ni++] = id;
    curCost += la[best];
    leaf[best] = b.left; lb[best] = ocLdBox(boxes + b.left); la[best] = halfArea(lb[best]);
    leaf[nl] = b.right; lb[nl] = ocLdBox(boxes + b.right); la[nl] = halfArea(lb[nl]);
    nl++;
  }

  // boxes of the subsets of the low 3 and the high 4 treelet leaves; any subset's box is the union of one of each
  Aabb empty;
  for (int k = 0; k < 3; k++) { empty.lo[k] = INFINITY; empty.hi[k] = -INFINITY; }
  Aabb lowB[8], highB[16];
  lowB[0] = empty; highB[0] = empty;
  for (int s = 1; s < 8; s++) {
    const int b = 31 - ocClz32((uint32_t)s);
    lowB[s] = unionOf(lowB[s ^ (1 << b)], lb[b]);
  }
  for (int s = 1; s < 16; s++) {
    const int b = 31 - ocClz32((uint32_t)s);
    highB[s] = unionOf(highB[s ^ (1 << b)], lb[3 + b]);
  }

  // dynamic programme: copt[s] = area(s) + min over splits (copt[p] + copt[s \ p]); single leaves cost nothing (their
  // subtrees are untouched). Subsets are visited in numeric order: every proper subset of s is smaller than s.
  float copt[OC_TREELET_SUBSETS];
  uint8_t part[OC_TREELET_SUBSETS];
  for (int s = 1; s < OC_TREELET_SUBSETS; s++) {
    if ((s & (s - 1)) == 0) { copt[s] = 0.f; part[s] = 0; continue; }
    const int low = s & -s, rest = s ^ low;
    float best = INFINITY;
    int bp = rest;
    for (int p = rest; p; p = (p - 1) & rest) { // p never holds the lowest leaf: each split is seen once
      const float c = copt[p] + copt[s ^ p];
      if (c < best) { best = c; bp = p; }
    }
    copt[s] = halfArea(unionOf(lowB[s & 7], highB[s >> 3])) + best;
    part[s] = (uint8_t)bp;
  }
  const float rootArea = halfArea(unionOf(lowB[7], highB[15]));
  const float newCost = copt[OC_TREELET_SUBSETS - 1] - rootArea;
  if (!(newCost < curCost * 0.9999f)) return false;

  // leaf counts of the treelet leaves (the rewritten nodes keep last - first + 1 = leaves below)
  uint32_t lsize[OC_TREELET_LEAVES];
  for (int k = 0; k < NL; k++) {
    if (leaf[k] >= leafBase) lsize[k] = 1u;
    else { const BinNode b = ocLdNode(nodes + leaf[k]); lsize[k] = b.last - b.first + 1u; }
  }
  // rebuild top-down, handing out the 6 internal ids again
  uint8_t stS[OC_TREELET_LEAVES];
  uint32_t stId[OC_TREELET_LEAVES];
  int sp = 0, used = 1;
  stS[sp] = (uint8_t)(OC_TREELET_SUBSETS - 1); stId[sp] = root; sp++;
  while (sp) {
    sp--;
    const int s = stS[sp];
    const uint32_t id = stId[sp];
    const int half[2] = {part[s], s ^ part[s]};
    uint32_t child[2];
    for (int h = 0; h < 2; h++) {
      const int c = half[h];
      if ((c & (c - 1)) == 0) child[h] = leaf[31 - ocClz32((uint32_t)c)];
      else {
        child[h] = inner[used++];
        stS[sp] = (uint8_t)c; stId[sp] = child[h]; sp++;
      }
      parent[child[h]] = id;
    }
    uint32_t sz = 0;
    for (int k = 0; k < NL; k++)
      if ((s >> k) & 1) sz += lsize[k];
    BinNode b;
    b.left = child[0]; b.right = child[1]; b.first = 0u; b.last = sz - 1u;
    nodes[id] = b;
    if (id != root) boxes[id] = unionOf(lowB[s & 7], highB[s >> 3]);
  }
  return true;
}

} // namespace oc
