// GPU BVH builder: Morton-code LBVH (Karras 2012) refined by SAH treelet restructuring (Karras & Aila 2013, treelet.h),
// collapsed into 4-wide, 128-byte nodes.
//
// Stands in for Embree 4.3.3 rtcCommitScene (un-vendored) as called by
// modules/cpu/common/Group.cpp:129-142 (one BVH per group over all primitives of all geometric
// models; geomID = index in the group's geometry array, Group.cpp:25) and
// modules/cpu/common/World.cpp:167-171 (top level over instances; instID = array index, :32).
//
// Pipeline (all on the render stream):
//   1. emit leaf records (3 x float4: pre-transformed triangle v0/e1/e2, sphere, box) + AABBs
//   2. scene bounds (ordered-uint atomics), 63-bit Morton keys of centroids
//   3. cub::DeviceRadixSort on (key, record index)
//   4. Karras hierarchy, bottom-up refit with arrival counters
//   4b. OC_TREELET_PASSES bottom-up passes of 7-leaf treelet restructuring (optimal topology by dynamic programming
//      over leaf subsets, surface-area cost)
//   5. level-synchronous collapse binary -> BVH4: children with the largest surface area are
//      opened first; subtrees of <= OC_LEAF_MAX (2) records become leaves; the pass hands every subtree its position
//      in depth-first leaf order
//   6. gather records into leaf order
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "runtime.h"
#include "shading.cuh"
#include "treelet.h"

namespace oc {

#define OC_TREELET_PASSES 2 // treelet restructuring passes over the LBVH (OSPCUDA_TREELET_PASSES overrides; 0 = plain LBVH)
#define OC_LEAF_MAX 2u // records per leaf: 2 beats 4 on every workload (C3 136 -> 106 ms, cornell 2.62 -> 2.41 ms; tools/sweep_leaf_max.sh)

__device__ __forceinline__ uint32_t encF(float f)
{
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b ^ 0x80000000u);
}
__host__ __device__ __forceinline__ float decF(uint32_t e)
{
  const uint32_t b = (e & 0x80000000u) ? (e ^ 0x80000000u) : ~e;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f; memcpy(&f, &b, 4); return f;
#endif
}

// position + radius of a curve segment (control points first .. first+3 of vertex.position_radius) at parameter t, with
// the basis polynomials of geometry/Curves.ispc:40-75 and every operation rounded separately (oracle: same expression)
__device__ __forceinline__ vec4f curveEval(const OcGeometry *g, uint32_t first, float t)
{
  const vec4f v0 = get4f(g->sphere, first), v1 = get4f(g->sphere, first + 1);
  const float s = __fsub_rn(1.0f, t);
  if (g->curveBasis == OC_CURVE_BASIS_LINEAR)
    return make_vec4f(__fadd_rn(__fmul_rn(s, v0.x), __fmul_rn(t, v1.x)), __fadd_rn(__fmul_rn(s, v0.y), __fmul_rn(t, v1.y)),
        __fadd_rn(__fmul_rn(s, v0.z), __fmul_rn(t, v1.z)), __fadd_rn(__fmul_rn(s, v0.w), __fmul_rn(t, v1.w)));
  const vec4f v2 = get4f(g->sphere, first + 2), v3 = get4f(g->sphere, first + 3);
  const float s2 = __fmul_rn(s, s), t2 = __fmul_rn(t, t), s2t = __fmul_rn(s2, t), t2s = __fmul_rn(t2, s);
  const float s3 = __fmul_rn(s2, s), t3 = __fmul_rn(t2, t);
  float b0, b1, b2, b3, cf;
  if (g->curveBasis == OC_CURVE_BASIS_BEZIER) {
    b0 = s3; b1 = __fmul_rn(3.0f, s2t); b2 = __fmul_rn(3.0f, t2s); b3 = t3; cf = 1.0f;
  } else if (g->curveBasis == OC_CURVE_BASIS_BSPLINE) {
    b0 = s3; b3 = t3; cf = 1.0f / 6.0f;
    b1 = __fadd_rn(__fadd_rn(__fmul_rn(4.0f, s3), t3), __fadd_rn(__fmul_rn(12.0f, s2t), __fmul_rn(6.0f, t2s)));
    b2 = __fadd_rn(__fadd_rn(__fmul_rn(4.0f, t3), s3), __fadd_rn(__fmul_rn(12.0f, t2s), __fmul_rn(6.0f, s2t)));
  } else {
    b0 = -s2t; b1 = __fadd_rn(2.0f, __fmul_rn(t2, __fsub_rn(__fmul_rn(3.0f, t), 5.0f)));
    b2 = __fadd_rn(2.0f, __fmul_rn(s2, __fsub_rn(__fmul_rn(3.0f, s), 5.0f))); b3 = -t2s; cf = 0.5f;
  }
  auto comb = [&](float a, float b, float c, float d) {
    return __fmul_rn(cf, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(b0, a), __fmul_rn(b1, b)), __fmul_rn(b2, c)), __fmul_rn(b3, d)));
  };
  return make_vec4f(comb(v0.x, v1.x, v2.x, v3.x), comb(v0.y, v1.y, v2.y, v3.y), comb(v0.z, v1.z, v2.z, v3.z), comb(v0.w, v1.w, v2.w, v3.w));
}

// ---- 1. records ---------------------------------------------------------------------------
__global__ void emitRecordsKernel(const OcModel *models, uint32_t numModels, const uint32_t *recOffset /*numModels+1*/,
    uint32_t numRecords, float4 *records, Aabb *boxes, uint32_t *sceneBounds /*6 enc*/)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  Aabb b;
  b.lo[0] = b.lo[1] = b.lo[2] = INFINITY;
  b.hi[0] = b.hi[1] = b.hi[2] = -INFINITY;
  float cullPad = 0.f;
  if (r < numRecords) {
    // model lookup: binary search in the prefix array
    uint32_t lo = 0, hi = numModels;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (recOffset[mid] <= r) lo = mid; else hi = mid;
    }
    const uint32_t geomID = lo;
    const OcGeometry *g = models[geomID].geom;
    const uint32_t local = r - recOffset[geomID];
    float4 r0, r1, r2;
    auto ext = [&](vec3f p) {
      b.lo[0] = fminf(b.lo[0], p.x); b.lo[1] = fminf(b.lo[1], p.y); b.lo[2] = fminf(b.lo[2], p.z);
      b.hi[0] = fmaxf(b.hi[0], p.x); b.hi[1] = fmaxf(b.hi[1], p.y); b.hi[2] = fmaxf(b.hi[2], p.z);
    };
    if (g->type == OC_GEOM_TRIANGLES || g->type == OC_GEOM_QUADS) {
      uint32_t primID, kind = OC_PRIM_TRI;
      vec3f v0, v1, v2;
      if (g->type == OC_GEOM_TRIANGLES) {
        primID = local;
        v0 = get3f(g->vertex, get1ui(g->index, primID, 0));
        v1 = get3f(g->vertex, get1ui(g->index, primID, 1));
        v2 = get3f(g->vertex, get1ui(g->index, primID, 2));
      } else {
        primID = local >> 1;
        const vec3f q0 = get3f(g->vertex, get1ui(g->index, primID, 0)), q1 = get3f(g->vertex, get1ui(g->index, primID, 1)),
                    q2 = get3f(g->vertex, get1ui(g->index, primID, 2)), q3 = get3f(g->vertex, get1ui(g->index, primID, 3));
        if ((local & 1u) == 0) { v0 = q0; v1 = q1; v2 = q3; }
        else { v0 = q2; v1 = q3; v2 = q1; kind = OC_PRIM_TRI_QUAD2ND; }
      }
      ext(v0); ext(v1); ext(v2);
      const vec3f e1 = make_vec3f(__fsub_rn(v0.x, v1.x), __fsub_rn(v0.y, v1.y), __fsub_rn(v0.z, v1.z));
      const vec3f e2 = make_vec3f(__fsub_rn(v2.x, v0.x), __fsub_rn(v2.y, v0.y), __fsub_rn(v2.z, v0.z));
      r0 = make_float4(v0.x, v0.y, v0.z, __uint_as_float(primID));
      r1 = make_float4(e1.x, e1.y, e1.z, __uint_as_float(geomID));
      r2 = make_float4(e2.x, e2.y, e2.z, __uint_as_float(kind));
    } else if (g->type == OC_GEOM_CURVES) {
      // one capsule of the tessellated curve segment: end points = basis at k/8 and (k+1)/8, radius = mean end radius
      const uint32_t primID = local / OC_CURVE_SUBDIV, k = local % OC_CURVE_SUBDIV;
      const uint32_t first = get1ui(g->index, primID, 0);
      const float u0 = (float)k * (1.f / OC_CURVE_SUBDIV), u1 = (float)(k + 1) * (1.f / OC_CURVE_SUBDIV);
      const vec4f a4 = curveEval(g, first, u0), b4 = curveEval(g, first, u1);
      const float rad = fmaxf(0.f, __fmul_rn(0.5f, __fadd_rn(a4.w, b4.w)));
      const vec3f pa = make_vec3f(a4.x, a4.y, a4.z), pb = make_vec3f(b4.x, b4.y, b4.z);
      // reported bounds: end points -+ radius, one rounding per operation (pyoracle reproduces them)
      ext(make_vec3f(__fsub_rn(fminf(pa.x, pb.x), rad), __fsub_rn(fminf(pa.y, pb.y), rad), __fsub_rn(fminf(pa.z, pb.z), rad)));
      ext(make_vec3f(__fadd_rn(fmaxf(pa.x, pb.x), rad), __fadd_rn(fmaxf(pa.y, pb.y), rad), __fadd_rn(fmaxf(pa.z, pb.z), rad)));
      // culling box of the BVH: a few ulp wider (the capsule test rounds every operation)
      cullPad = rad * 1e-3f + 4e-6f * fmaxf(fmaxf(reduce_max(vabs(pa)), reduce_max(vabs(pb))), rad);
      r0 = make_float4(pa.x, pa.y, pa.z, __uint_as_float(primID));
      r1 = make_float4(pb.x, pb.y, pb.z, __uint_as_float(geomID));
      r2 = make_float4(rad, u0, 1.f / OC_CURVE_SUBDIV, __uint_as_float((uint32_t)OC_PRIM_CAPSULE));
    } else if (g->type == OC_GEOM_SPHERES) {
      const vec4f s = get4f(g->sphere, local);
      ext(make_vec3f(s.x - s.w, s.y - s.w, s.z - s.w));
      ext(make_vec3f(s.x + s.w, s.y + s.w, s.z + s.w));
      r0 = make_float4(s.x, s.y, s.z, __uint_as_float(local));
      r1 = make_float4(s.w, 0.f, 0.f, __uint_as_float(geomID));
      r2 = make_float4(0.f, 0.f, 0.f, __uint_as_float((uint32_t)OC_PRIM_SPHERE));
    } else {
      const float *p = reinterpret_cast<const float *>(g->box.addr + (int64_t)local * g->box.byteStride);
      const vec3f lo3 = make_vec3f(p[0], p[1], p[2]), hi3 = make_vec3f(p[3], p[4], p[5]);
      ext(lo3); ext(hi3);
      r0 = make_float4(lo3.x, lo3.y, lo3.z, __uint_as_float(local));
      r1 = make_float4(hi3.x, hi3.y, hi3.z, __uint_as_float(geomID));
      r2 = make_float4(0.f, 0.f, 0.f, __uint_as_float((uint32_t)OC_PRIM_BOX));
    }
    records[3 * (size_t)r + 0] = r0; records[3 * (size_t)r + 1] = r1; records[3 * (size_t)r + 2] = r2;
    Aabb bp = b;
    for (int a = 0; a < 3; a++) { bp.lo[a] -= cullPad; bp.hi[a] += cullPad; }
    boxes[r] = bp;
  }
  // block-level bounds reduction, then 6 atomics per warp
#pragma unroll
  for (int a = 0; a < 3; a++) {
    float lo = b.lo[a], hi = b.hi[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
      atomicMin(sceneBounds + a, encF(lo));
      atomicMax(sceneBounds + 3 + a, encF(hi));
    }
  }
}

__global__ void boundsOfBoxesKernel(const Aabb *boxes, uint32_t n, uint32_t *sceneBounds)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    float lo = r < n ? boxes[r].lo[a] : INFINITY, hi = r < n ? boxes[r].hi[a] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
      atomicMin(sceneBounds + a, encF(lo));
      atomicMax(sceneBounds + 3 + a, encF(hi));
    }
  }
}

// ---- 2. Morton keys -------------------------------------------------------------------------
__device__ __forceinline__ uint64_t expandBits21(uint64_t v)
{
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
// `bits` per axis (<= 21): fewer bits put more primitives into one cell; the hierarchy over equal keys is split by index,
// i.e. balanced, which bounds its depth (the fall-back build of a group whose first hierarchy came out too deep)
__global__ void mortonKernel(const Aabb *boxes, uint32_t n, const uint32_t *sceneBounds, uint64_t *keys, uint32_t *ids, uint32_t bits)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t key = 0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float lo = decF(sceneBounds[a]), hi = decF(sceneBounds[3 + a]);
    const float c = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);
    const float ext = hi - lo;
    float f = ext > 0.f ? (c - lo) / ext : 0.f;
    const float cells = (float)(1u << bits);
    f = fminf(fmaxf(f * cells, 0.f), cells - 1.f);
    key |= expandBits21((uint64_t)f << (21u - bits)) << (2 - a);
  }
  keys[i] = key;
  ids[i] = i;
}

// ---- 4. Karras hierarchy ----------------------------------------------------------------------
// node ids: internal i in [0, n-1), leaf j as (n-1)+j (BinNode, karrasNode: treelet.h)
__global__ void karrasKernel(const uint64_t *keys, int n, BinNode *nodes, uint32_t *parent)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const BinNode b = karrasNode(keys, n, i);
  nodes[i] = b;
  parent[b.left] = i;
  parent[b.right] = i;
  if (i == 0) parent[0] = 0xFFFFFFFFu;
}

__global__ void refitKernel(const uint32_t *sortedIds, const Aabb *primBoxes, int n, const BinNode *nodes, const uint32_t *parent,
    Aabb *nodeBoxes /*2n-1*/, uint32_t *flags)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t cur = n - 1 + j;
  nodeBoxes[cur] = primBoxes[sortedIds[j]];
  __threadfence();
  uint32_t p = (n > 1) ? parent[cur] : 0xFFFFFFFFu;
  while (p != 0xFFFFFFFFu) {
    if (atomicAdd(flags + p, 1u) == 0) return; // first arrival: sibling not done yet
    __threadfence();
    // children were written by other threads: read through L2 (ld.cg), not a stale L1 line
    const float *pa = reinterpret_cast<const float *>(nodeBoxes + nodes[p].left);
    const float *pb = reinterpret_cast<const float *>(nodeBoxes + nodes[p].right);
    Aabb u;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      u.lo[k] = fminf(__ldcg(pa + k), __ldcg(pb + k));
      u.hi[k] = fmaxf(__ldcg(pa + 3 + k), __ldcg(pb + 3 + k));
    }
    nodeBoxes[p] = u;
    __threadfence();
    p = parent[p];
  }
}

// ---- 4b. treelet restructuring (SAH refinement of the LBVH; treelet.h) ---------------------------------------------
// One thread per leaf climbs the tree; the second thread to reach a node owns the finished subtree below it (same
// arrival protocol as refitKernel) and restructures the 7-leaf treelet rooted there. Rewrites stay inside that subtree,
// whose other threads have all retired; reads of nodes / boxes / parents written by them go through L2.
__global__ void treeletKernel(int n, BinNode *nodes, uint32_t *parent, Aabb *nodeBoxes, uint32_t *flags, uint32_t *rewritten)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t p = __ldcg(parent + (n - 1 + j));
  uint32_t count = 0;
  while (p != 0xFFFFFFFFu) {
    if (atomicAdd(flags + p, 1u) == 0) break;
    __threadfence();
    const BinNode b = ocLdNode(nodes + p);
    if (b.last - b.first + 1u >= (uint32_t)OC_TREELET_LEAVES) {
      count += optimizeTreelet(p, n, nodes, parent, nodeBoxes) ? 1u : 0u;
      __threadfence();
    }
    p = __ldcg(parent + p);
  }
  if (count) atomicAdd(rewritten, count);
}

// surface-area cost of the binary tree (sum of internal-node half areas / root half area), OSPCUDA_BVH_STATS=1 only
__global__ void sahCostKernel(const Aabb *nodeBoxes, uint32_t numBin, double *sum)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double a = i < numBin ? (double)halfArea(nodeBoxes[i]) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0 && a != 0.0) atomicAdd(sum, a);
}

// ---- 5. collapse ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t subtreeSize(const BinNode *nodes, int n, uint32_t id)
{
  return id >= (uint32_t)(n - 1) ? 1u : nodes[id].last - nodes[id].first + 1u;
}
// `first`: position of the subtree's first record in leaf order = number of leaves to its left in a depth-first walk.
// It is handed down here (the treelet pass reorders subtrees, so the sorted-key ranges of the Karras pass no longer
// describe them); leaves write leafOrder[position] = record index, which gatherRecordsKernel then follows.
// `pushed`: stack entries a depth-first traversal holds when it arrives at this node in the worst case (every sibling on
// the path passed the slab test) = sum over the wide ancestors of (children - 1)
struct WorkItem { uint32_t bin, wide, first, pushed; };

__global__ void collapseKernel(const BinNode *nodes, const Aabb *nodeBoxes, int n, const WorkItem *inQ, uint32_t inCount,
    WorkItem *outQ, uint32_t *outCount, uint32_t *wideCount, OcNode4 *wide, uint32_t leafMax, const uint32_t *sortedIds,
    uint32_t *leafOrder, uint32_t *stackNeed)
{
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= inCount) return;
  const WorkItem item = inQ[w];
  uint32_t slot[4], first[4];
  int cnt = 0;
  if (subtreeSize(nodes, n, item.bin) <= leafMax) {
    slot[0] = item.bin; first[0] = item.first; cnt = 1; // tiny tree: root holds one leaf
  } else {
    slot[0] = nodes[item.bin].left; first[0] = item.first;
    slot[1] = nodes[item.bin].right; first[1] = item.first + subtreeSize(nodes, n, slot[0]);
    cnt = 2;
    while (cnt < 4) {
      int best = -1;
      float bestA = -1.f;
      for (int k = 0; k < cnt; k++) {
        if (subtreeSize(nodes, n, slot[k]) <= leafMax) continue;
        const float a = halfArea(nodeBoxes[slot[k]]);
        if (a > bestA) { bestA = a; best = k; }
      }
      if (best < 0) break;
      const uint32_t open = slot[best];
      slot[best] = nodes[open].left;
      slot[cnt] = nodes[open].right;
      first[cnt] = first[best] + subtreeSize(nodes, n, slot[best]);
      cnt++;
    }
  }
  const uint32_t pushedBelow = item.pushed + (uint32_t)cnt - 1u;
  atomicMax(stackNeed, pushedBelow);
  OcNode4 out;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (k < cnt) {
      const Aabb b = nodeBoxes[slot[k]];
      out.lox[k] = b.lo[0]; out.loy[k] = b.lo[1]; out.loz[k] = b.lo[2];
      out.hix[k] = b.hi[0]; out.hiy[k] = b.hi[1]; out.hiz[k] = b.hi[2];
      const uint32_t sz = subtreeSize(nodes, n, slot[k]);
      if (sz <= leafMax) {
        out.child[k] = 0x80000000u | (sz << 27) | first[k];
        // records of the leaf in depth-first order (leafMax <= 15)
        uint32_t stk[16], sp = 0, pos = first[k];
        stk[sp++] = slot[k];
        while (sp) {
          const uint32_t id = stk[--sp];
          if (id >= (uint32_t)(n - 1)) leafOrder[pos++] = sortedIds[id - (uint32_t)(n - 1)];
          else { stk[sp++] = nodes[id].right; stk[sp++] = nodes[id].left; }
        }
      } else {
        const uint32_t wi = atomicAdd(wideCount, 1u);
        out.child[k] = wi;
        const uint32_t qi = atomicAdd(outCount, 1u);
        WorkItem ni; ni.bin = slot[k]; ni.wide = wi; ni.first = first[k]; ni.pushed = pushedBelow;
        outQ[qi] = ni;
      }
    } else {
      out.lox[k] = out.loy[k] = out.loz[k] = INFINITY;
      out.hix[k] = out.hiy[k] = out.hiz[k] = -INFINITY;
      out.child[k] = 0xFFFFFFFFu;
    }
    out.pad[k] = 0;
  }
  // 8 x 128-bit stores
  const float4 *src = reinterpret_cast<const float4 *>(&out);
  float4 *dst = reinterpret_cast<float4 *>(wide + item.wide);
#pragma unroll
  for (int k = 0; k < 8; k++) dst[k] = src[k];
}

__global__ void gatherRecordsKernel(const float4 *records, const uint32_t *leafOrder, uint32_t n, float4 *out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t s = 3 * (size_t)leafOrder[i];
  out[3 * (size_t)i + 0] = records[s + 0];
  out[3 * (size_t)i + 1] = records[s + 1];
  out[3 * (size_t)i + 2] = records[s + 2];
}

template <typename T>
struct DevBuf
{
  T *p = nullptr;
  cudaError_t alloc(size_t n) { return n ? cudaMalloc((void **)&p, n * sizeof(T)) : cudaSuccess; }
  ~DevBuf() { if (p) cudaFree(p); }
};

static uint32_t g_optMortonBits = 0, g_optLeafMax = 0; // ospcuda_bvh_set_build_options (0 = default)
static uint32_t g_lastStackNeed = 0; // of the most recent build (ospcuda_bvh_last_stack_need)
static uint32_t g_peakStackNeed = 0; // largest since the last reset (ospcuda_bvh_peak_stack_need)

// shared tail: boxes (+records) -> OcBVH
static int buildFromBoxes(Aabb *d_boxes, float4 *d_records, uint32_t n, const uint32_t *d_sceneBounds, OcBVH *bvh)
{
  cudaStream_t st = (cudaStream_t)ospcuda_stream();
  const uint32_t B = 256, G = (n + B - 1) / B;
  DevBuf<uint64_t> keys, keysSorted;
  DevBuf<uint32_t> ids, idsSorted, parent, flags, counters, leafOrder;
  DevBuf<BinNode> bin;
  DevBuf<Aabb> nodeBoxes;
  DevBuf<WorkItem> q0, q1;
  DevBuf<OcNode4> wideTmp;
  DevBuf<uint8_t> cubTmp;
  OC_TRY(keys.alloc(n)); OC_TRY(keysSorted.alloc(n)); OC_TRY(ids.alloc(n)); OC_TRY(idsSorted.alloc(n));
  mortonKernel<<<G, B, 0, st>>>(d_boxes, n, d_sceneBounds, keys.p, ids.p, g_optMortonBits ? std::min(21u, g_optMortonBits) : 21u);
  oc_count_launch();
  size_t tmpBytes = 0;
  OC_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys.p, keysSorted.p, ids.p, idsSorted.p, (int)n, 0, 63, st));
  OC_TRY(cubTmp.alloc(tmpBytes));
  OC_TRY(cub::DeviceRadixSort::SortPairs(cubTmp.p, tmpBytes, keys.p, keysSorted.p, ids.p, idsSorted.p, (int)n, 0, 63, st));
  oc_count_launch();

  const uint32_t numBin = n > 1 ? n - 1 : 0;
  OC_TRY(bin.alloc(numBin ? numBin : 1)); OC_TRY(parent.alloc(2 * (size_t)n)); OC_TRY(flags.alloc(numBin ? numBin : 1));
  OC_TRY(nodeBoxes.alloc(2 * (size_t)n));
  OC_TRY(cudaMemsetAsync(flags.p, 0, sizeof(uint32_t) * (numBin ? numBin : 1), st));
  if (numBin) {
    karrasKernel<<<(numBin + B - 1) / B, B, 0, st>>>(keysSorted.p, (int)n, bin.p, parent.p);
    oc_count_launch();
  }
  refitKernel<<<G, B, 0, st>>>(idsSorted.p, d_boxes, (int)n, bin.p, parent.p, nodeBoxes.p, flags.p);
  oc_count_launch();
  // SAH refinement: treelet restructuring passes (each pass lowers the tree's surface-area cost; the third adds little)
  static const int treeletPasses = getenv("OSPCUDA_TREELET_PASSES") ? std::max(0, atoi(getenv("OSPCUDA_TREELET_PASSES"))) : OC_TREELET_PASSES;
  static const bool stats = getenv("OSPCUDA_BVH_STATS") && atoi(getenv("OSPCUDA_BVH_STATS")) > 0;
  DevBuf<double> costSum;
  auto reportCost = [&](const char *what) -> int {
    if (!stats || !numBin) return 0;
    if (!costSum.p) OC_TRY(costSum.alloc(1));
    OC_TRY(cudaMemsetAsync(costSum.p, 0, sizeof(double), st));
    sahCostKernel<<<(numBin + B - 1) / B, B, 0, st>>>(nodeBoxes.p, numBin, costSum.p);
    double h_sum = 0.0; Aabb rootBox; uint32_t h_rew = 0;
    OC_TRY(cudaMemcpyAsync(&h_sum, costSum.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    OC_TRY(cudaMemcpyAsync(&rootBox, nodeBoxes.p, sizeof(Aabb), cudaMemcpyDeviceToHost, st));
    OC_TRY(cudaMemcpyAsync(&h_rew, counters.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    OC_TRY(cudaStreamSynchronize(st));
    fprintf(stderr, "[ospcuda bvh] n=%u %s: SAH cost (internal area / root area) = %.3f, treelets rewritten so far = %u\n", n, what,
        h_sum / (double)halfArea(rootBox), h_rew);
    return 0;
  };
  OC_TRY(counters.alloc(4));
  OC_TRY(cudaMemsetAsync(counters.p + 2, 0, 2 * sizeof(uint32_t), st)); // [2] treelets rewritten, [3] traversal stack need
  if (n >= OC_TREELET_LEAVES && treeletPasses > 0) {
    if (reportCost("LBVH")) return -1;
    for (int pass = 0; pass < treeletPasses; pass++) {
      OC_TRY(cudaMemsetAsync(flags.p, 0, sizeof(uint32_t) * numBin, st));
      treeletKernel<<<G, B, 0, st>>>((int)n, bin.p, parent.p, nodeBoxes.p, flags.p, counters.p + 2);
      oc_count_launch();
      if (reportCost("after a treelet pass")) return -1;
    }
  }

  // collapse, level by level
  const uint32_t maxWide = numBin ? numBin : 1;
  OC_TRY(wideTmp.alloc(maxWide)); OC_TRY(q0.alloc(maxWide)); OC_TRY(q1.alloc(maxWide)); OC_TRY(leafOrder.alloc(n));
  uint32_t h_counts[2] = {0, 1}; // [0] next-queue size, [1] wide nodes allocated
  OC_TRY(cudaMemcpyAsync(counters.p, h_counts, sizeof(h_counts), cudaMemcpyHostToDevice, st));
  WorkItem root; root.bin = 0u /* n == 1: the single leaf's id is (n-1)+0 = 0 too */; root.wide = 0; root.first = 0; root.pushed = 0;
  OC_TRY(cudaMemcpyAsync(q0.p, &root, sizeof(root), cudaMemcpyHostToDevice, st));
  uint32_t inCount = 1;
  WorkItem *inQ = q0.p, *outQ = q1.p;
  static const uint32_t leafMaxDefault = getenv("OSPCUDA_LEAF_MAX") ? std::min(15u, std::max(1u, (uint32_t)atoi(getenv("OSPCUDA_LEAF_MAX")))) : OC_LEAF_MAX;
  const uint32_t leafMax = g_optLeafMax ? std::min(15u, g_optLeafMax) : leafMaxDefault;
  while (inCount) {
    collapseKernel<<<(inCount + 127) / 128, 128, 0, st>>>(bin.p, nodeBoxes.p, (int)n, inQ, inCount, outQ, counters.p, counters.p + 1,
        wideTmp.p, leafMax, idsSorted.p, leafOrder.p, counters.p + 3);
    oc_count_launch();
    OC_TRY(cudaMemcpyAsync(h_counts, counters.p, sizeof(h_counts), cudaMemcpyDeviceToHost, st));
    OC_TRY(cudaStreamSynchronize(st));
    inCount = h_counts[0];
    const uint32_t zero = 0;
    OC_TRY(cudaMemcpyAsync(counters.p, &zero, sizeof(zero), cudaMemcpyHostToDevice, st));
    WorkItem *t = inQ; inQ = outQ; outQ = t;
  }
  const uint32_t numWide = h_counts[1];

  OcNode4 *nodesOut = nullptr;
  OC_TRY(cudaMalloc((void **)&nodesOut, sizeof(OcNode4) * (size_t)numWide));
  OC_TRY(cudaMemcpyAsync(nodesOut, wideTmp.p, sizeof(OcNode4) * (size_t)numWide, cudaMemcpyDeviceToDevice, st));
  float4 *recOut = nullptr;
  OC_TRY(cudaMalloc((void **)&recOut, sizeof(float4) * 3 * (size_t)n));
  gatherRecordsKernel<<<G, B, 0, st>>>(d_records, leafOrder.p, n, recOut);
  oc_count_launch();
  uint32_t h_enc[6];
  OC_TRY(cudaMemcpyAsync(h_enc, d_sceneBounds, sizeof(h_enc), cudaMemcpyDeviceToHost, st));
  OC_TRY(cudaMemcpyAsync(&g_lastStackNeed, counters.p + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  OC_TRY(cudaStreamSynchronize(st));
  if (g_lastStackNeed > g_peakStackNeed) g_peakStackNeed = g_lastStackNeed;
  bvh->nodes = nodesOut;
  bvh->prims = reinterpret_cast<const OcVec4f *>(recOut);
  bvh->numNodes = numWide;
  bvh->numRecords = n;
  for (int k = 0; k < 6; k++) bvh->bounds[k] = decF(h_enc[k]);
  return oc_check(cudaGetLastError(), "bvh build");
}

static void emptyBVH(OcBVH *bvh)
{
  bvh->nodes = nullptr; bvh->prims = nullptr; bvh->numNodes = 0; bvh->numRecords = 0;
  for (int k = 0; k < 3; k++) { bvh->bounds[k] = INFINITY; bvh->bounds[3 + k] = -INFINITY; }
}

__global__ void initBoundsKernel(uint32_t *b)
{
  if (threadIdx.x < 3) b[threadIdx.x] = 0xFFFFFFFFu;
  else if (threadIdx.x < 6) b[threadIdx.x] = 0u;
}

__global__ void instanceRecordsKernel(const Aabb *boxes, uint32_t n, float4 *records)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  records[3 * (size_t)i + 0] = make_float4(boxes[i].lo[0], boxes[i].lo[1], boxes[i].lo[2], __uint_as_float(i));
  records[3 * (size_t)i + 1] = make_float4(boxes[i].hi[0], boxes[i].hi[1], boxes[i].hi[2], 0.f);
  records[3 * (size_t)i + 2] = make_float4(0.f, 0.f, 0.f, __uint_as_float((uint32_t)OC_PRIM_INSTANCE));
}

} // namespace oc

using namespace oc;

extern "C" int ospcuda_bvh_build_group(const OcModel *d_models, uint32_t numModels, const uint32_t *h_numPrims,
    const uint32_t *h_types, OcBVH *bvh, float *buildMs)
{
  emptyBVH(bvh);
  g_lastStackNeed = 0;
  if (buildMs) *buildMs = 0.f;
  std::vector<uint32_t> off(numModels + 1, 0);
  uint64_t total = 0;
  for (uint32_t i = 0; i < numModels; i++) {
    off[i] = (uint32_t)total;
    total += (uint64_t)h_numPrims[i] * (h_types[i] == OC_GEOM_QUADS ? 2u : (h_types[i] == OC_GEOM_CURVES ? OC_CURVE_SUBDIV : 1u));
  }
  off[numModels] = (uint32_t)total;
  if (total == 0) return 0;
  if (total >= (1ull << 27)) return oc_set_error(cudaErrorInvalidValue, "ospcuda_bvh_build_group: more than 2^27 leaf records");
  const uint32_t n = (uint32_t)total;
  cudaStream_t st = (cudaStream_t)ospcuda_stream();
  cudaEvent_t e0, e1;
  OC_TRY(cudaEventCreate(&e0)); OC_TRY(cudaEventCreate(&e1));
  OC_TRY(cudaEventRecord(e0, st));
  DevBuf<uint32_t> d_off, d_bounds;
  DevBuf<float4> records;
  DevBuf<Aabb> boxes;
  OC_TRY(d_off.alloc(numModels + 1)); OC_TRY(d_bounds.alloc(6)); OC_TRY(records.alloc(3 * (size_t)n)); OC_TRY(boxes.alloc(n));
  OC_TRY(cudaMemcpyAsync(d_off.p, off.data(), sizeof(uint32_t) * (numModels + 1), cudaMemcpyHostToDevice, st));
  initBoundsKernel<<<1, 32, 0, st>>>(d_bounds.p);
  emitRecordsKernel<<<(n + 255) / 256, 256, 0, st>>>(d_models, numModels, d_off.p, n, records.p, boxes.p, d_bounds.p);
  oc_count_launch();
  // every group gets a hierarchy, however small (one record: a root node with one leaf): traversal is one state machine
  // over nodes, without a second code path for record lists
  const int rc = buildFromBoxes(boxes.p, records.p, n, d_bounds.p, bvh);
  OC_TRY(cudaEventRecord(e1, st));
  OC_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (buildMs) *buildMs = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

extern "C" int ospcuda_bvh_build_tlas(const float *h_bounds, uint32_t numInstances, OcBVH *bvh)
{
  emptyBVH(bvh);
  g_lastStackNeed = 0;
  if (numInstances == 0) return 0;
  cudaStream_t st = (cudaStream_t)ospcuda_stream();
  DevBuf<Aabb> boxes;
  DevBuf<float4> records;
  DevBuf<uint32_t> d_bounds;
  OC_TRY(boxes.alloc(numInstances)); OC_TRY(records.alloc(3 * (size_t)numInstances)); OC_TRY(d_bounds.alloc(6));
  OC_TRY(cudaMemcpyAsync(boxes.p, h_bounds, sizeof(Aabb) * numInstances, cudaMemcpyHostToDevice, st));
  initBoundsKernel<<<1, 32, 0, st>>>(d_bounds.p);
  boundsOfBoxesKernel<<<(numInstances + 255) / 256, 256, 0, st>>>(boxes.p, numInstances, d_bounds.p);
  instanceRecordsKernel<<<(numInstances + 255) / 256, 256, 0, st>>>(boxes.p, numInstances, records.p);
  oc_count_launch();
  return buildFromBoxes(boxes.p, records.p, numInstances, d_bounds.p, bvh);
}

extern "C" uint32_t ospcuda_bvh_last_stack_need(void) { return g_lastStackNeed; }
extern "C" int ospcuda_bvh_set_build_options(uint32_t mortonBitsPerAxis, uint32_t leafMax)
{
  g_optMortonBits = mortonBitsPerAxis;
  g_optLeafMax = leafMax;
  return 0;
}
extern "C" uint32_t ospcuda_bvh_peak_stack_need(int reset)
{
  const uint32_t v = g_peakStackNeed;
  if (reset) g_peakStackNeed = 0;
  return v;
}

extern "C" int ospcuda_bvh_free(OcBVH *bvh)
{
  if (bvh->nodes) cudaFree((void *)bvh->nodes);
  if (bvh->prims) cudaFree((void *)bvh->prims);
  emptyBVH(bvh);
  return 0;
}

// ---- structured volume macrocells ------------------------------------------------------------------
namespace oc {
// value range of the (up to) 17^3 vertices spanned by one macrocell of 16^3 cells
__global__ void macrocellKernel(const float *voxels, uint32_t dx, uint32_t dy, uint32_t dz, float *ranges, uint32_t mx, uint32_t my,
    uint32_t mz)
{
  const uint32_t mc = blockIdx.x;
  const uint32_t cx = mc % mx, cy = (mc / mx) % my, cz = mc / (mx * my);
  float lo = INFINITY, hi = -INFINITY;
  for (uint32_t t = threadIdx.x; t < 17 * 17 * 17; t += blockDim.x) {
    const uint32_t x = cx * 16 + t % 17, y = cy * 16 + (t / 17) % 17, z = cz * 16 + t / 289;
    if (x < dx && y < dy && z < dz) {
      const float v = voxels[x + (size_t)dx * (y + (size_t)dy * z)];
      if (!isnan(v)) { lo = fminf(lo, v); hi = fmaxf(hi, v); }
    }
  }
  __shared__ float slo[8], shi[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
    ranges[2 * (size_t)mc] = lo;
    ranges[2 * (size_t)mc + 1] = hi;
  }
}
} // namespace oc

extern "C" int ospcuda_volume_macrocells(const float *d_voxels, const uint32_t dims[3], float *d_ranges, const uint32_t mcDims[3])
{
  const uint32_t total = mcDims[0] * mcDims[1] * mcDims[2];
  if (total == 0) return 0;
  macrocellKernel<<<total, 256, 0, (cudaStream_t)ospcuda_stream()>>>(d_voxels, dims[0], dims[1], dims[2], d_ranges, mcDims[0], mcDims[1],
      mcDims[2]);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_volume_macrocells");
}
