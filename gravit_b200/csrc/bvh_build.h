// bvh_build.h -- per-thread bodies of the GPU BLAS build:
//   tri_setup   : triangle AABB + 30-bit Morton key of its centroid           (one thread per face)
//   [radix sort by key, stable -- kernels.cu]
//   leaf_setup  : sorted leaf boxes                                            (one thread per sorted face)
//   karras_node : binary radix tree over the sorted keys (Karras, HPG 2012)    (one thread per internal node)
//   refit_up    : bottom-up AABB fit with one atomic flag per internal node    (one thread per leaf)
//   collapse    : binary tree -> 8-wide compressed nodes (128 bytes, 7-bit quantised child boxes), surface-area-greedy
//                 child expansion, octant-ordered slots; triangles re-laid per wide node  (one thread per wide node,
//                 level by level through a task queue)
// Replaces Embree's rtcNewTriangleMesh + rtcCommit (adapter/embree/EmbreeMeshAdapter.cpp:125-162).
#pragma once
#include "gvt_device.h"

namespace gvt {

// a value written by another thread of the same launch: read through L2
#if defined(__CUDA_ARCH__)
#define GVT_LDCG1(p) __ldcg(p)
#else
#define GVT_LDCG1(p) (*(p))
#endif

struct BuildCtx {
  const float *verts;    // 3 per vertex
  const uint32_t *faces; // 3 per face, 0-based
  unsigned n;            // faces
  V3 mlo, mext;          // mesh bbox min and extent (hi - lo)
  float pad;             // absolute padding of leaf boxes (conservative traversal)
  // sort
  uint32_t *keys, *vals; // Morton key and face index (sorted in place by the radix sort)
  // binary tree: node ids 0..n-2 internal, n-1..2n-2 leaves (leaf k = sorted position k)
  float4 *bn_lo, *bn_hi;
  int2 *bn_child;
  int *bn_parent;
  int2 *bn_range;
  unsigned *bn_flag;
  // wide output
  float4 *wnodes; // GVT_NODE_WORDS per node
  float4 *wtris;  // 3 per triangle record slot (GVT_LEAF_MAX slots per leaf child)
  unsigned *node_count, *tri_count, *tri_real; // wide nodes, triangle record slots, triangles referenced
  unsigned wnode_cap;
  // insertion-based optimisation (reinsert_*): per binary node
  int4 *ri_pick;                  // {target y, pivot, gain bits, 0}; y < 0: no improving move
  unsigned long long *ri_lock;    // highest (gain, node) key that wants to touch this node in the current iteration
  unsigned *ri_go;                // 1: the node's move is carried out in this iteration
  unsigned *ri_cnt;               // triangles below an internal node (relayout)
  // optimal widening (collapse_dp_node): per internal binary node, null = surface-area-greedy widening
  float *dp_cost;         // [7 * node + i - 1] = C(node, i), i = 1..7
  unsigned char *dp_pick; // [8 * node + j - 1], j = 2..8: left share k of the best distribution of j roots; [8 * node]: bit 0 =
                          // "C(node,1) is a leaf", bits 1..6 = "C(node,i) is C(node,i-1)" for i = 2..7
};

// absolute padding of the leaf boxes: the canonical triangle test is evaluated in fp32, so a ray can be accepted while
// passing a few ulp (of the ray-origin-to-triangle distance) outside the exact triangle; the boxes must not cull it
GVT_HD float blas_leaf_pad(V3 mext) { return 4e-6f * (mext.x + mext.y + mext.z) + 1e-30f; }

GVT_HD uint32_t expand_bits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

GVT_HD void tri_verts(const BuildCtx &c, unsigned f, V3 &a, V3 &b, V3 &cc) {
  unsigned i0 = c.faces[3 * f], i1 = c.faces[3 * f + 1], i2 = c.faces[3 * f + 2];
  a = v3(c.verts[3 * i0], c.verts[3 * i0 + 1], c.verts[3 * i0 + 2]);
  b = v3(c.verts[3 * i1], c.verts[3 * i1 + 1], c.verts[3 * i1 + 2]);
  cc = v3(c.verts[3 * i2], c.verts[3 * i2 + 1], c.verts[3 * i2 + 2]);
}

// Morton key: centroid of the triangle's AABB (0.5f*lo + 0.5f*hi), normalised by the mesh bbox with an IEEE
// division, scaled by 1024, clamped to [0,1023], truncated; x in the top bit of each triple.  Bit-exact with
// oracle/gvt_oracle.cpp gvto_morton.
GVT_HD void tri_setup(const BuildCtx &c, unsigned f) {
  V3 a, b, d;
  tri_verts(c, f, a, b, d);
  V3 lo = v3(fminf(fminf(a.x, b.x), d.x), fminf(fminf(a.y, b.y), d.y), fminf(fminf(a.z, b.z), d.z));
  V3 hi = v3(fmaxf(fmaxf(a.x, b.x), d.x), fmaxf(fmaxf(a.y, b.y), d.y), fmaxf(fmaxf(a.z, b.z), d.z));
  V3 cen = xadd(xmul(lo, 0.5f), xmul(hi, 0.5f));
  float nx = c.mext.x > 0.f ? xdiv(xsub(cen.x, c.mlo.x), c.mext.x) : 0.f;
  float ny = c.mext.y > 0.f ? xdiv(xsub(cen.y, c.mlo.y), c.mext.y) : 0.f;
  float nz = c.mext.z > 0.f ? xdiv(xsub(cen.z, c.mlo.z), c.mext.z) : 0.f;
  uint32_t qx = (uint32_t)fminf(fmaxf(xmul(nx, 1024.f), 0.f), 1023.f);
  uint32_t qy = (uint32_t)fminf(fmaxf(xmul(ny, 1024.f), 0.f), 1023.f);
  uint32_t qz = (uint32_t)fminf(fmaxf(xmul(nz, 1024.f), 0.f), 1023.f);
  c.keys[f] = (expand_bits10(qx) << 2) | (expand_bits10(qy) << 1) | expand_bits10(qz);
  c.vals[f] = f;
}

// leaf k (sorted position) -> padded AABB
GVT_HD void leaf_setup(const BuildCtx &c, unsigned k) {
  V3 a, b, d;
  tri_verts(c, c.vals[k], a, b, d);
  float p = c.pad;
  unsigned id = c.n - 1 + k;
  c.bn_lo[id] = make_float4(fminf(fminf(a.x, b.x), d.x) - p, fminf(fminf(a.y, b.y), d.y) - p, fminf(fminf(a.z, b.z), d.z) - p, 0.f);
  c.bn_hi[id] = make_float4(fmaxf(fmaxf(a.x, b.x), d.x) + p, fmaxf(fmaxf(a.y, b.y), d.y) + p, fmaxf(fmaxf(a.z, b.z), d.z) + p, 0.f);
}

// common-prefix length of sorted keys i and j; ties broken by position so that all keys are distinct
GVT_HD int karras_delta(const BuildCtx &c, int i, int j) {
  if (j < 0 || j >= (int)c.n) return -1;
  uint32_t a = c.keys[i], b = c.keys[j];
  if (a != b) return clz32(a ^ b);
  return 32 + clz32((uint32_t)i ^ (uint32_t)j);
}

GVT_HD void karras_node(const BuildCtx &c, int i) {
  const int n = (int)c.n;
  int d = (karras_delta(c, i, i + 1) - karras_delta(c, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = karras_delta(c, i, i - d);
  int lmax = 2;
  while (karras_delta(c, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (karras_delta(c, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = karras_delta(c, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (karras_delta(c, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + (d < 0 ? d : 0);
  int lo = i < j ? i : j, hi = i < j ? j : i;
  int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
  int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
  c.bn_child[i] = make_int2(left, right);
  c.bn_range[i] = make_int2(lo, hi);
  c.bn_parent[left] = i;
  c.bn_parent[right] = i;
  if (i == 0) c.bn_parent[0] = -1;
}

// ---- binned-SAH binary tree (top-down, level-synchronous) ----------------------------------------------------
// The Morton order only seeds the element order (locality); the tree itself is built top-down: per level, every active
// node bins the centroids of its triangles (GVT_SAH_BINS bins on each axis, bin bounds grown with integer atomics),
// picks the split with the smallest SA(L)*|L| + SA(R)*|R|, and its element range is stably partitioned (one device-wide
// scan of the goes-left flags per level).  Nodes of at most GVT_SAH_SMALL triangles evaluate every centroid threshold
// exactly instead of binning.  The output has the layout the Karras tree had (internal nodes 0..n-2 with root 0, leaf
// n-1+k = sorted position k, bn_child / bn_range / bn_parent / bn_lo / bn_hi), so refit is not needed and the wide
// collapse below is unchanged.  Replaces the SAH builder inside Embree's rtcCommit (EmbreeMeshAdapter.cpp:125-162).
#ifndef GVT_SAH_BINS
#define GVT_SAH_BINS 16
#endif
#ifndef GVT_SAH_SMALL
#define GVT_SAH_SMALL 8
#endif

GVT_HD int fenc(float f) { // order-preserving float -> signed int
  int i = (int)f2u(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
GVT_HD float fdec(int i) { return u2f((uint32_t)(i >= 0 ? i : i ^ 0x7FFFFFFF)); }

struct SahBin {
  int lo[3], hi[3]; // encoded box of the triangles whose centroid falls into the bin
  int cnt, pad;
};
struct SahNode {   // an active node of the current level
  int lo, hi;      // element range [lo, hi)
  int id;          // binary node id
  int bslot;       // index of its 3 * GVT_SAH_BINS bins, -1 for small nodes
  int cb[6];       // encoded centroid bounds (min.xyz, max.xyz)
  // split decision
  int axis, bin, mode; // mode 0: left = (bin index <= bin); 1: left = first half by position; 2: left = (centroid <= thr)
  float thr;
  int n_left, child_active[2];
};
GVT_HD int sah_bin(float c, float cmin, float cmax) {
  float ext = cmax - cmin;
  if (!(ext > 0.f)) return 0;
  int b = (int)((c - cmin) * ((float)GVT_SAH_BINS * (1.f - 1e-6f) / ext));
  return b < 0 ? 0 : (b >= GVT_SAH_BINS ? GVT_SAH_BINS - 1 : b);
}
GVT_HD float sah_area(const float *lo, const float *hi) {
  float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return dx * dy + dy * dz + dz * dx;
}
// choose the split of a binned node; also returns the node box (union of its bins)
GVT_HD void sah_eval_bins(const SahBin *bins /* [3][GVT_SAH_BINS] */, SahNode &nd, float *blo, float *bhi) {
  float best = FLT_MAX;
  nd.mode = 1, nd.axis = 0, nd.bin = 0;
  for (int q = 0; q < 3; ++q) blo[q] = FLT_MAX, bhi[q] = -FLT_MAX;
  for (int a = 0; a < 3; ++a) {
    const SahBin *B = bins + a * GVT_SAH_BINS;
    float ra[GVT_SAH_BINS];
    int rc[GVT_SAH_BINS];
    float l[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, h[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    int n = 0;
    for (int b = GVT_SAH_BINS - 1; b >= 0; --b) {
      if (B[b].cnt) {
        n += B[b].cnt;
        for (int q = 0; q < 3; ++q) l[q] = fminf(l[q], fdec(B[b].lo[q])), h[q] = fmaxf(h[q], fdec(B[b].hi[q]));
      }
      rc[b] = n, ra[b] = n ? sah_area(l, h) : 0.f;
    }
    if (a == 0)
      for (int q = 0; q < 3; ++q) blo[q] = l[q], bhi[q] = h[q];
    for (int q = 0; q < 3; ++q) l[q] = FLT_MAX, h[q] = -FLT_MAX;
    n = 0;
    for (int b = 0; b + 1 < GVT_SAH_BINS; ++b) {
      if (B[b].cnt) {
        n += B[b].cnt;
        for (int q = 0; q < 3; ++q) l[q] = fminf(l[q], fdec(B[b].lo[q])), h[q] = fmaxf(h[q], fdec(B[b].hi[q]));
      }
      if (!n || !rc[b + 1]) continue;
      float cost = sah_area(l, h) * (float)n + ra[b + 1] * (float)rc[b + 1];
      if (cost < best) best = cost, nd.mode = 0, nd.axis = a, nd.bin = b, nd.n_left = n;
    }
  }
}
// small node: every centroid of the node is a threshold candidate on every axis (exact sweep, O(k^2))
GVT_HD void sah_eval_small(const float4 *tb_lo, const float4 *tb_hi, const uint32_t *idx, SahNode &nd, float *blo, float *bhi) {
  float best = FLT_MAX;
  nd.mode = 1, nd.axis = 0, nd.bin = 0, nd.thr = 0.f;
  const int k = nd.hi - nd.lo;
  for (int q = 0; q < 3; ++q) blo[q] = FLT_MAX, bhi[q] = -FLT_MAX;
  for (int i = 0; i < k; ++i) {
    float4 l = tb_lo[idx[nd.lo + i]], h = tb_hi[idx[nd.lo + i]];
    blo[0] = fminf(blo[0], l.x), blo[1] = fminf(blo[1], l.y), blo[2] = fminf(blo[2], l.z);
    bhi[0] = fmaxf(bhi[0], h.x), bhi[1] = fmaxf(bhi[1], h.y), bhi[2] = fmaxf(bhi[2], h.z);
  }
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < k; ++i) {
      float4 li = tb_lo[idx[nd.lo + i]], hi_ = tb_hi[idx[nd.lo + i]];
      float thr = 0.5f * ((&li.x)[a] + (&hi_.x)[a]);
      float l0[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, h0[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX }, l1[3] = { FLT_MAX, FLT_MAX, FLT_MAX },
            h1[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
      int n0 = 0, n1 = 0;
      for (int j = 0; j < k; ++j) {
        float4 l = tb_lo[idx[nd.lo + j]], h = tb_hi[idx[nd.lo + j]];
        float c = 0.5f * ((&l.x)[a] + (&h.x)[a]);
        if (c <= thr) {
          n0++;
          l0[0] = fminf(l0[0], l.x), l0[1] = fminf(l0[1], l.y), l0[2] = fminf(l0[2], l.z);
          h0[0] = fmaxf(h0[0], h.x), h0[1] = fmaxf(h0[1], h.y), h0[2] = fmaxf(h0[2], h.z);
        } else {
          n1++;
          l1[0] = fminf(l1[0], l.x), l1[1] = fminf(l1[1], l.y), l1[2] = fminf(l1[2], l.z);
          h1[0] = fmaxf(h1[0], h.x), h1[1] = fmaxf(h1[1], h.y), h1[2] = fmaxf(h1[2], h.z);
        }
      }
      if (!n0 || !n1) continue;
      float cost = sah_area(l0, h0) * (float)n0 + sah_area(l1, h1) * (float)n1;
      if (cost < best) best = cost, nd.mode = 2, nd.axis = a, nd.thr = thr, nd.n_left = n0;
    }
}
// does element `f` at position `pos` go to the left child?
GVT_HD bool sah_goes_left(const SahNode &nd, const float4 *tb_lo, const float4 *tb_hi, uint32_t f, int pos) {
  if (nd.mode == 1) return (pos - nd.lo) < (nd.hi - nd.lo) / 2;
  float c = 0.5f * ((&tb_lo[f].x)[nd.axis] + (&tb_hi[f].x)[nd.axis]);
  if (nd.mode == 2) return c <= nd.thr;
  return sah_bin(c, fdec(nd.cb[nd.axis]), fdec(nd.cb[3 + nd.axis])) <= nd.bin;
}
GVT_HD void tri_box_padded(const BuildCtx &c, unsigned f, float4 &lo, float4 &hi) {
  V3 a, b, d;
  tri_verts(c, f, a, b, d);
  float p = c.pad;
  lo = make_float4(fminf(fminf(a.x, b.x), d.x) - p, fminf(fminf(a.y, b.y), d.y) - p, fminf(fminf(a.z, b.z), d.z) - p, 0.f);
  hi = make_float4(fmaxf(fmaxf(a.x, b.x), d.x) + p, fmaxf(fmaxf(a.y, b.y), d.y) + p, fmaxf(fmaxf(a.z, b.z), d.z) + p, 0.f);
}

// ---- wide-node collapse ------------------------------------------------------------------------------------
// Wide node = 128 bytes = one cache line = eight 16-byte words (coalesced LDG.128 each):
//   u0: frame origin.xyz (fp32), {ex, ey, ez, imask} bytes   e* = biased exponent of 128 * grid step; imask bit s = slot s
//                                                            holds an internal child
//   u1: child_base, tri_base, tmask, 0                       tmask bit 8k+s = slot s is a leaf with more than k triangles
//   u2: x lo-planes, u3: x hi-planes, u4/u5: y, u6/u7: z     word j of a block = field(slot j) << 16 | field(slot j+4)
// A plane is a 7-bit grid coordinate q stored as the 16-bit field 0x3F80 | q -- the upper half of the float 1 + q/128 --
// so that the traversal gets the value of slot j for free (the word itself, read as a float; the lower field adds a known
// 0.249 quantum) and the value of slot j+4 with one shift: no integer-to-float conversion, no byte extraction
// (traverse.h node_step).  Internal children of a node are stored consecutively from child_base in slot order; the
// triangles of its leaf children (at most GVT_LEAF_MAX each) are stored from tri_base, three record slots per leaf in slot
// order.  Compressed wide BVH after Ylitie, Karras, Laine (HPG 2017), re-laid for sm_100's ALU/FMA pipe balance.
// Triangle record = 48 bytes: {v0.xyz, prim id}, {e1.xyz, 0}, {e2.xyz, 0} with e1 = v1 - v0, e2 = v2 - v0.
#define GVT_LEAF_MAX 3 /* triangles per leaf child */
#define GVT_NODE_WORDS 8 /* 16-byte words per wide node */
#define GVT_QMAX 127

GVT_HD float box_area(float4 lo, float4 hi) {
  float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
GVT_HD int bnode_count(const BuildCtx &c, int b) {
  if (b >= (int)c.n - 1) return 1;
  int2 r = c.bn_range[b];
  return r.y - r.x + 1;
}
GVT_HD int bnode_first(const BuildCtx &c, int b) {
  if (b >= (int)c.n - 1) return b - ((int)c.n - 1);
  return c.bn_range[b].x;
}

#define GVT_QPAD 0.02f /* the node frame starts this many grid steps below the children's minimum, so that the conservative
                          rounding margins of the planes are never eaten by the clamp to the grid */
// biased exponent e such that (GVT_QMAX - 0.2) * 2^(e-127) >= ext (smallest such), clamped to [1,247]
GVT_HD uint32_t quant_exponent(float ext) {
  if (!(ext > 0.f)) return 1u;
  float s = ext * (1.0f / ((float)GVT_QMAX - 0.2f)) * 1.0001f; // room for GVT_QPAD and the margins below GVT_QMAX
  uint32_t u = f2u(s);
  uint32_t e = (u >> 23) & 0xFFu;
  if (u & 0x7FFFFFu) e += 1u; // round up to the next power of two
  if (e < 1u) e = 1u;
  if (e > 247u) e = 247u;
  return e;
}

// ---- insertion-based optimisation of the binary tree -----------------------------------------------------------
// Top-down SAH decides every split once, with what it knows at that moment; on architectural meshes (big walls next to
// fine detail) a node often ends up in a subtree where it drags boxes open.  Bittner, Hapala and Havran (2013) repair
// that by taking nodes out and re-inserting them where the surface-area cost grows least; Meister and Bittner (2018)
// give the parallel form used here.  One iteration:
//   reinsert_search  every node x looks for the node y that it would best become the sibling of.  Walking up from x,
//                    each ancestor ("pivot") contributes the subtree on its other side as a search region; d is the
//                    area freed so far: x's parent p disappears (its sibling takes its place) and every ancestor between
//                    p and the pivot shrinks to the box of what is left below it.  Inside a region the cost of reaching y
//                    is the growth of every box on the way down plus the new parent box A(x u y); branches that cannot
//                    beat the best gain so far are cut.
//   reinsert_lock    a move rewrites pointers of six nodes (x, p, its sibling, its grandparent, y, y's parent); each move
//                    bids for them with atomicMax of (gain, x).
//   reinsert_check   a move that holds all six is a candidate;
//   reinsert_confirm it is dropped if a node on its way up to the pivot moves as well (two such moves could carry each
//                    other's subtree: a cycle);
//   reinsert_apply   the confirmed moves rewire their pointers; nobody else reads or writes them in that phase.  Boxes
//                    are re-fitted bottom-up afterwards (k_refit).  Every phase only reads what the previous one wrote, so
//                    the outcome does not depend on thread order.
// Moves change the left-to-right order of the leaves, which the widening relies on (a leaf slot is a contiguous run of
// sorted positions), so the last step renumbers the leaves in depth-first order (relayout_*).  Hits do not depend on the
// tree; Morton keys are taken before.
#define GVT_RI_STACK 96
// iterations for a mesh of n faces: GVTB200_REINSERT overrides; by default small and medium meshes get 8 (a search is
// cheap there and architectural models gain 8-14 % fewer node visits), very large ones none (uniform tessellations such as
// an isosurface gain nothing and every iteration is a pass over the whole tree)
static inline int reinsert_iterations(size_t n) {
  const char *e = getenv("GVTB200_REINSERT");
  if (e) return atoi(e) < 0 ? 0 : atoi(e);
  return n <= 4000000 ? 8 : 0;
}
GVT_HD float box_area4(float4 lo, float4 hi) {
  float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
GVT_HD float union_area4(float4 al, float4 ah, float4 bl, float4 bh) {
  float dx = fmaxf(ah.x, bh.x) - fminf(al.x, bl.x), dy = fmaxf(ah.y, bh.y) - fminf(al.y, bl.y), dz = fmaxf(ah.z, bh.z) - fminf(al.z, bl.z);
  return dx * dy + dy * dz + dz * dx;
}
GVT_HD int bn_sibling(const BuildCtx &c, int parent, int child) {
  const int2 k = c.bn_child[parent];
  return k.x == child ? k.y : k.x;
}
// explores one region (the subtree of `root`) for the best sibling of x; d = area freed if x leaves its place
GVT_HD void reinsert_region(const BuildCtx &c, int root, int pivot, int skip, float4 xl, float4 xh, float ax, float d,
                            float &best_gain, int &best_y, int &best_pivot) {
  int sn[GVT_RI_STACK];
  float si[GVT_RI_STACK];
  int sp = 0;
  sn[sp] = root, si[sp] = 0.f, sp++;
  const int first_leaf = (int)c.n - 1;
  while (sp) {
    --sp;
    const int y = sn[sp];
    const float ind = si[sp];
    const float4 yl = c.bn_lo[y], yh = c.bn_hi[y];
    const float direct = union_area4(xl, xh, yl, yh);
    const float gain = d - direct - ind;
    if (gain > best_gain && y != skip) best_gain = gain, best_y = y, best_pivot = pivot;
    if (y < first_leaf) {
      const float ind2 = ind + direct - box_area4(yl, yh); // y grows to x u y when x goes somewhere below it
      if (d - ax - ind2 > best_gain && sp + 2 <= GVT_RI_STACK) {
        const int2 k = c.bn_child[y];
        sn[sp] = k.x, si[sp] = ind2, sp++;
        sn[sp] = k.y, si[sp] = ind2, sp++;
      }
    }
  }
}
GVT_HD void reinsert_search(const BuildCtx &c, int x) {
  int4 pick = make_int4(-1, -1, 0, 0);
  const int p = c.bn_parent[x];
  const int g = p >= 0 ? c.bn_parent[p] : -1;
  if (g >= 0) { // the root and its children stay where they are
    const float4 xl = c.bn_lo[x], xh = c.bn_hi[x];
    const float ax = box_area4(xl, xh);
    const int s = bn_sibling(c, p, x);
    float d = box_area4(c.bn_lo[p], c.bn_hi[p]);
    float best_gain = d * 1e-4f; // a move must pay for itself
    int best_y = -1, best_pivot = -1;
    reinsert_region(c, s, p, s, xl, xh, ax, d, best_gain, best_y, best_pivot); // below the sibling (the sibling itself = no move)
    float4 sl = c.bn_lo[s], sh = c.bn_hi[s]; // box of what is left below the pivot's child on x's side
    int cur = p, piv = g;
    while (piv >= 0) {
      const int other = bn_sibling(c, piv, cur);
      reinsert_region(c, other, piv, -1, xl, xh, ax, d, best_gain, best_y, best_pivot);
      const float4 ol = c.bn_lo[other], oh = c.bn_hi[other];
      sl = make_float4(fminf(sl.x, ol.x), fminf(sl.y, ol.y), fminf(sl.z, ol.z), 0.f);
      sh = make_float4(fmaxf(sh.x, oh.x), fmaxf(sh.y, oh.y), fmaxf(sh.z, oh.z), 0.f);
      d += box_area4(c.bn_lo[piv], c.bn_hi[piv]) - box_area4(sl, sh); // above here the pivot loses x as well
      cur = piv, piv = c.bn_parent[piv];
    }
    if (best_y >= 0) pick = make_int4(best_y, best_pivot, (int)f2u(best_gain), 0);
  }
  c.ri_pick[x] = pick;
}
GVT_HD unsigned long long reinsert_key(int4 pick, int x) { return ((unsigned long long)(uint32_t)pick.z << 32) | (uint32_t)x; }
#if defined(__CUDA_ARCH__)
#define GVT_ATOMIC_MAX64(p, v) atomicMax((p), (v))
#else
#define GVT_ATOMIC_MAX64(p, v) (*(p) = *(p) < (v) ? (v) : *(p))
#endif
// the nodes whose pointers a move of x rewrites: x, its parent p, its sibling s, its grandparent g, the target y and y's parent q
GVT_HD void reinsert_nodes(const BuildCtx &c, int x, int4 pick, int out[6]) {
  const int p = c.bn_parent[x];
  out[0] = x, out[1] = p, out[2] = bn_sibling(c, p, x), out[3] = c.bn_parent[p], out[4] = pick.x, out[5] = c.bn_parent[pick.x];
}
GVT_HD void reinsert_lock(const BuildCtx &c, int x) {
  const int4 pick = c.ri_pick[x];
  if (pick.x < 0) return;
  const unsigned long long key = reinsert_key(pick, x);
  int nd[6];
  reinsert_nodes(c, x, pick, nd);
  for (int i = 0; i < 6; ++i) GVT_ATOMIC_MAX64(&c.ri_lock[nd[i]], key);
}
// after every bid is in: does x hold all six nodes?  (the answer goes into pick.w; the tree is still untouched)
GVT_HD void reinsert_check(const BuildCtx &c, int x) {
  const int4 pick = c.ri_pick[x];
  if (pick.x < 0) return;
  const unsigned long long key = reinsert_key(pick, x);
  int nd[6];
  reinsert_nodes(c, x, pick, nd);
  bool mine = true;
  for (int i = 0; i < 6; ++i) mine = mine && c.ri_lock[nd[i]] == key;
  c.ri_pick[x].w = mine ? 1 : 0;
}
// Two moves with disjoint node sets can still carry each other's subtree (x goes below x'' while x'' goes below x: a cycle cut
// off from the root).  A move is therefore dropped when any node on its way -- from p up to the pivot and from y up to
// the pivot -- is itself a node that moves in this iteration.  Reads only the verdicts of reinsert_check, writes ri_go.
GVT_HD void reinsert_confirm(const BuildCtx &c, int x) {
  const int4 pick = c.ri_pick[x];
  unsigned go = 0;
  if (pick.x >= 0 && pick.w == 1) {
    go = 1;
    int steps = 0;
    for (int a = c.bn_parent[x]; a != pick.y && go; a = c.bn_parent[a])
      if (a < 0 || ++steps > 4096 || c.ri_pick[a].w == 1) go = 0;
    for (int a = pick.x; a != pick.y && go; a = c.bn_parent[a])
      if (a < 0 || ++steps > 8192 || c.ri_pick[a].w == 1) go = 0;
    if (go && c.ri_pick[pick.y].w == 1) go = 0;
  }
  c.ri_go[x] = go;
}
GVT_HD bool reinsert_apply(const BuildCtx &c, int x) {
  if (!c.ri_go[x]) return false;
  const int4 pick = c.ri_pick[x];
  const int y = pick.x, p = c.bn_parent[x], g = c.bn_parent[p], s = bn_sibling(c, p, x);
  // take p out: its sibling-side child moves up
  if (c.bn_child[g].x == p) c.bn_child[g].x = s; else c.bn_child[g].y = s;
  c.bn_parent[s] = g;
  // put p in above y
  const int q = c.bn_parent[y];
  if (c.bn_child[q].x == y) c.bn_child[q].x = p; else c.bn_child[q].y = p;
  c.bn_parent[p] = q;
  c.bn_child[p] = make_int2(y, x);
  c.bn_parent[y] = p;
  return true;
}
// triangles below internal node b (bottom-up: both children are done)
GVT_HD void relayout_count_node(const BuildCtx &c, int b) {
  const int2 k = c.bn_child[b];
  const int first_leaf = (int)c.n - 1;
  c.ri_cnt[b] = (k.x >= first_leaf ? 1u : GVT_LDCG1(&c.ri_cnt[k.x])) + (k.y >= first_leaf ? 1u : GVT_LDCG1(&c.ri_cnt[k.y]));
}
// relayout: sorted position of a leaf / first position of an internal node = triangles to the left of it in depth-first
// order = sum, over the ancestors entered through their right child, of the left child's triangle count
GVT_HD unsigned relayout_first(const BuildCtx &c, int node) {
  unsigned first = 0;
  const int first_leaf = (int)c.n - 1;
  int cur = node;
  for (int a = c.bn_parent[cur]; a >= 0; cur = a, a = c.bn_parent[a]) {
    const int2 k = c.bn_child[a];
    if (k.y == cur) first += k.x >= first_leaf ? 1u : c.ri_cnt[k.x];
  }
  return first;
}

// ---- optimal widening (Ylitie, Karras, Laine, HPG 2017, section 3.1) -------------------------------------------
// C(n, i) = cheapest SAH cost of binary subtree n represented by at most i roots (wide nodes or leaves):
//   C(n, 1) = A_n * P_n * C_prim if P_n <= GVT_LEAF_MAX (a leaf slot), else C_dist(n, 8) + A_n * C_node (a wide node)
//   C(n, i) = min( C_dist(n, i), C(n, i - 1) )                      i = 2..7
//   C_dist(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k)
// evaluated bottom-up (one thread climbs from every leaf, the second arrival at a node does its row), then followed
// top-down by collapse_gather_dp.  C_node : C_prim is the measured cost ratio of a node visit (8 L1 wavefronts, ~130
// instructions) to a triangle test (3 wavefronts, ~40 instructions).
#define GVT_DP_CNODE 1.0f
#ifndef GVT_DP_CPRIM
#define GVT_DP_CPRIM 0.35f
#endif

GVT_HD float dp_cost_of(const BuildCtx &c, int b, int i) { // C(b, i), i = 1..7 (more roots than 7 never help a leaf triangle)
  if (b >= (int)c.n - 1) return box_area(c.bn_lo[b], c.bn_hi[b]) * GVT_DP_CPRIM;
  return GVT_LDCG1(&c.dp_cost[7 * (size_t)b + (i > 7 ? 7 : i) - 1]); // written by another thread of the same launch: through L2
}
GVT_HD void collapse_dp_node(const BuildCtx &c, int b) {
  const int2 ch = c.bn_child[b];
  float cl[8], cr[8];
  for (int i = 1; i <= 7; ++i) cl[i] = dp_cost_of(c, ch.x, i), cr[i] = dp_cost_of(c, ch.y, i);
  float dist[9];
  unsigned char *pick = c.dp_pick + 8 * (size_t)b;
  for (int j = 2; j <= 8; ++j) {
    float best = FLT_MAX;
    int bk = 1;
    for (int k = 1; k < j; ++k) {
      float v = cl[k] + cr[j - k];
      if (v < best) best = v, bk = k;
    }
    dist[j] = best;
    pick[j - 1] = (unsigned char)bk;
  }
  const float area = box_area(c.bn_lo[b], c.bn_hi[b]);
  const int cnt = bnode_count(c, b);
  float c1 = dist[8] + area * GVT_DP_CNODE;
  unsigned char flags = 0;
  if (cnt <= GVT_LEAF_MAX) c1 = area * (float)cnt * GVT_DP_CPRIM, flags |= 1u; // as one root it is a leaf slot
  float *out = c.dp_cost + 7 * (size_t)b;
  out[0] = c1;
  for (int i = 2; i <= 7; ++i) {
    float v = dist[i];
    if (out[i - 2] <= v) v = out[i - 2], flags |= (unsigned char)(1u << (i - 1));
    out[i - 1] = v;
  }
  pick[0] = flags;
}
// children of the wide node rooted at binary node `root`, by the tables: expand (node, roots allowed) pairs, left first
GVT_HD int collapse_gather_dp(const BuildCtx &c, int root, int *ch) {
  int nch = 0, sp = 0;
  int sn[16], sj[16];
  sn[sp] = root, sj[sp] = 8, sp++;
  while (sp) {
    --sp;
    const int b = sn[sp];
    int j = sj[sp];
    if (b >= (int)c.n - 1) { // a single triangle
      ch[nch++] = b;
      continue;
    }
    const unsigned char flags = c.dp_pick[8 * (size_t)b];
    if (j < 8) { // j == 8 is the root of this wide node: it is a wide node, distribute 8
      while (j >= 2 && ((flags >> (j - 1)) & 1u)) --j; // C(b, j) is C(b, j - 1)
      if (j == 1) {
        ch[nch++] = b;
        continue;
      }
    }
    const int k = c.dp_pick[8 * (size_t)b + j - 1];
    const int2 kid = c.bn_child[b];
    sn[sp] = kid.y, sj[sp] = j - k, sp++;
    sn[sp] = kid.x, sj[sp] = k, sp++;
  }
  return nch;
}

struct CollapseTask {
  int bnode; // binary node to turn into a wide node
  int wnode; // index of the wide node to write
};

struct WideNode { // what collapse_node decides, before the output ranges are known
  int slot_b[8];      // binary node per slot, -1 = empty
  int cnt[8];         // triangles of a leaf slot, 0 for internal / empty slots
  unsigned qlo[8][3]; // quantised planes
  unsigned qhi[8][3];
  uint32_t imask, tmask;
  int n_internal, n_leaves, n_tris;
  float4 w0;
};

// Builds wide node t.wnode from binary subtree t.bnode.  Host emulation and the device kernel share this body; the
// device kernel reserves the output ranges with atomicAdd.
GVT_HD void collapse_node(const BuildCtx &c, CollapseTask t, WideNode &w) {
  // 1. gather up to 8 children: by the optimal-widening tables, or by repeatedly opening the child with the largest
  //    surface area that holds more than GVT_LEAF_MAX triangles
  int ch[8];
  int nch = 0;
  if (t.bnode >= (int)c.n - 1) {
    ch[nch++] = t.bnode; // single-triangle mesh: the root is a leaf
  } else if (c.dp_cost) {
    nch = collapse_gather_dp(c, t.bnode, ch);
  } else {
    int2 k = c.bn_child[t.bnode];
    ch[nch++] = k.x;
    ch[nch++] = k.y;
    while (nch < 8) {
      int best = -1;
      float besta = -1.f;
      for (int i = 0; i < nch; ++i) {
        if (bnode_count(c, ch[i]) <= GVT_LEAF_MAX) continue;
        float a = box_area(c.bn_lo[ch[i]], c.bn_hi[ch[i]]);
        if (a > besta) besta = a, best = i;
      }
      if (best < 0) break;
      int2 k2 = c.bn_child[ch[best]];
      ch[best] = k2.x;
      ch[nch++] = k2.y;
    }
  }
  // 2. node frame
  float4 lo = c.bn_lo[ch[0]], hi = c.bn_hi[ch[0]];
  for (int i = 1; i < nch; ++i) {
    float4 l = c.bn_lo[ch[i]], h = c.bn_hi[ch[i]];
    lo.x = fminf(lo.x, l.x), lo.y = fminf(lo.y, l.y), lo.z = fminf(lo.z, l.z);
    hi.x = fmaxf(hi.x, h.x), hi.y = fmaxf(hi.y, h.y), hi.z = fmaxf(hi.z, h.z);
  }
  uint32_t ex = quant_exponent(hi.x - lo.x), ey = quant_exponent(hi.y - lo.y), ez = quant_exponent(hi.z - lo.z);
  float isx = 1.0f / u2f(ex << 23), isy = 1.0f / u2f(ey << 23), isz = 1.0f / u2f(ez << 23);
  lo.x -= GVT_QPAD * u2f(ex << 23), lo.y -= GVT_QPAD * u2f(ey << 23), lo.z -= GVT_QPAD * u2f(ez << 23);
  // 3. slot assignment: slot s bit2/bit1/bit0 = child sits towards +x/+y/+z of the node centre, so that
  //    visiting slots in descending (s ^ (7 - ray octant)) order is front to back.  Greedy max-cost matching.
  float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
  float cost[8][8];
  for (int i = 0; i < nch; ++i) {
    float4 l = c.bn_lo[ch[i]], h = c.bn_hi[ch[i]];
    float dx = 0.5f * (l.x + h.x) - cx, dy = 0.5f * (l.y + h.y) - cy, dz = 0.5f * (l.z + h.z) - cz;
    for (int s = 0; s < 8; ++s) cost[i][s] = ((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 1) ? dz : -dz);
  }
  for (int s = 0; s < 8; ++s) w.slot_b[s] = -1;
  bool used[8] = { false, false, false, false, false, false, false, false };
  for (int k = 0; k < nch; ++k) {
    int bi = -1, bs = -1;
    float bc = -FLT_MAX;
    for (int i = 0; i < nch; ++i) {
      if (used[i]) continue;
      for (int s = 0; s < 8; ++s) {
        if (w.slot_b[s] >= 0) continue;
        if (cost[i][s] > bc) bc = cost[i][s], bi = i, bs = s;
      }
    }
    used[bi] = true;
    w.slot_b[bs] = ch[bi];
  }
  // 4. masks + quantised boxes
  w.imask = 0, w.tmask = 0;
  w.n_internal = 0, w.n_leaves = 0, w.n_tris = 0;
  for (int s = 0; s < 8; ++s) {
    w.cnt[s] = 0;
    w.qlo[s][0] = w.qlo[s][1] = w.qlo[s][2] = GVT_QMAX; // empty slot: inverted box
    w.qhi[s][0] = w.qhi[s][1] = w.qhi[s][2] = 0u;
    int b = w.slot_b[s];
    if (b < 0) continue;
    int cnt = bnode_count(c, b);
    if (cnt > GVT_LEAF_MAX) {
      w.imask |= 1u << s;
      w.n_internal++;
    } else {
      w.cnt[s] = cnt;
      for (int k = 0; k < cnt; ++k) w.tmask |= 1u << (8 * k + s);
      w.n_leaves++;
      w.n_tris += cnt;
    }
    float4 l = c.bn_lo[b], h = c.bn_hi[b];
    // conservative: floor/ceil with a 4e-3 grid-step margin (covers the rounding of the traversal's dequantisation
    // constants, traverse.h node_step); the clamp to the grid never binds (GVT_QPAD, quant_exponent)
    const float qm = (float)GVT_QMAX;
    float fx = floorf((l.x - lo.x) * isx - 4e-3f), fy = floorf((l.y - lo.y) * isy - 4e-3f), fz = floorf((l.z - lo.z) * isz - 4e-3f);
    float gx = ceilf((h.x - lo.x) * isx + 4e-3f), gy = ceilf((h.y - lo.y) * isy + 4e-3f), gz = ceilf((h.z - lo.z) * isz + 4e-3f);
    w.qlo[s][0] = (unsigned)fminf(fmaxf(fx, 0.f), qm), w.qlo[s][1] = (unsigned)fminf(fmaxf(fy, 0.f), qm), w.qlo[s][2] = (unsigned)fminf(fmaxf(fz, 0.f), qm);
    w.qhi[s][0] = (unsigned)fminf(fmaxf(gx, 0.f), qm), w.qhi[s][1] = (unsigned)fminf(fmaxf(gy, 0.f), qm), w.qhi[s][2] = (unsigned)fminf(fmaxf(gz, 0.f), qm);
  }
  // the stored exponent is that of 128 grid steps (the traversal multiplies it with the float 1 + q/128)
  w.w0 = make_float4(lo.x, lo.y, lo.z, u2f((ex + 7u) | ((ey + 7u) << 8) | ((ez + 7u) << 16) | (w.imask << 24)));
}

GVT_HD uint32_t plane_word(const unsigned q[8][3], int j, int comp) {
  return ((0x3F80u | (q[j][comp] & 0x7Fu)) << 16) | (0x3F80u | (q[j + 4][comp] & 0x7Fu));
}

// writes the node words and this node's triangle records; child tasks go to out[0..n_internal)
GVT_HD void collapse_emit(const BuildCtx &c, CollapseTask t, const WideNode &w, unsigned child_base, unsigned tri_base, CollapseTask *out) {
  uint4 *nd = reinterpret_cast<uint4 *>(c.wnodes) + GVT_NODE_WORDS * (size_t)t.wnode;
  nd[0] = make_uint4(f2u(w.w0.x), f2u(w.w0.y), f2u(w.w0.z), f2u(w.w0.w));
  nd[1] = make_uint4(child_base, tri_base, w.tmask, 0u);
  for (int comp = 0; comp < 3; ++comp) {
    nd[2 + 2 * comp] = make_uint4(plane_word(w.qlo, 0, comp), plane_word(w.qlo, 1, comp), plane_word(w.qlo, 2, comp), plane_word(w.qlo, 3, comp));
    nd[3 + 2 * comp] = make_uint4(plane_word(w.qhi, 0, comp), plane_word(w.qhi, 1, comp), plane_word(w.qhi, 2, comp), plane_word(w.qhi, 3, comp));
  }
  int ki = 0, leaf_rank = 0;
  for (int s = 0; s < 8; ++s) {
    int b = w.slot_b[s];
    if (b < 0) continue;
    if (w.cnt[s] == 0) { // internal child -> next level task
      out[ki].bnode = b;
      out[ki].wnode = (int)child_base + ki;
      ki++;
    } else {
      int first = bnode_first(c, b);
      for (int k = 0; k < w.cnt[s]; ++k) {
        unsigned f = c.vals[first + k];
        V3 a, bb, d;
        tri_verts(c, f, a, bb, d);
        V3 e1 = xsub(bb, a), e2 = xsub(d, a);
        float4 *tr = c.wtris + 3 * (size_t)(tri_base + GVT_LEAF_MAX * leaf_rank + k);
        tr[0] = make_float4(a.x, a.y, a.z, u2f(f));
        tr[1] = make_float4(e1.x, e1.y, e1.z, 0.f);
        tr[2] = make_float4(e2.x, e2.y, e2.z, 0.f);
      }
      leaf_rank++;
    }
  }
}

} // namespace gvt
