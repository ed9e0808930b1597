// traverse.h -- per-thread traversal bodies: top-level instance routing, instance transform, 8-wide compressed
// BVH traversal (closest hit / any hit) and the canonical ray/triangle test.
#pragma once
#include "gvt_device.h"

namespace gvt {

// ---------------------------------------------------------------------------------------------------------
// Canonical ray/triangle test (Moeller-Trumbore).  Stands in for Embree 2.15's rtcIntersect8/rtcOccluded8
// (adapter/embree/EmbreeMeshAdapter.cpp:474,375); u weights v1, v weights v2 (:555-558).  The operation
// sequence -- which products are fused and which are not -- is part of the contract with the oracle
// (oracle/gvt_oracle.cpp triTest): identical t,u,v bits on both sides.  Valid hit: tnear < t < tfar.
// ---------------------------------------------------------------------------------------------------------
GVT_HD bool tri_test(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float tnear, float tfar, float &t, float &u, float &v) {
  float px = xfma(d.y, e2.z, -xmul(d.z, e2.y));
  float py = xfma(d.z, e2.x, -xmul(d.x, e2.z));
  float pz = xfma(d.x, e2.y, -xmul(d.y, e2.x));
  float det = xfma(e1.z, pz, xfma(e1.y, py, xmul(e1.x, px)));
  float inv = xdiv(1.0f, det);
  float sx = xsub(o.x, v0.x), sy = xsub(o.y, v0.y), sz = xsub(o.z, v0.z);
  float uu = xmul(xfma(sz, pz, xfma(sy, py, xmul(sx, px))), inv);
  if (!(uu >= 0.f && uu <= 1.f)) return false;
  float qx = xfma(sy, e1.z, -xmul(sz, e1.y));
  float qy = xfma(sz, e1.x, -xmul(sx, e1.z));
  float qz = xfma(sx, e1.y, -xmul(sy, e1.x));
  float vv = xmul(xfma(d.z, qz, xfma(d.y, qy, xmul(d.x, qx))), inv);
  if (!(vv >= 0.f && xadd(uu, vv) <= 1.f)) return false;
  float tt = xmul(xfma(e2.z, qz, xfma(e2.y, qy, xmul(e2.x, qx))), inv);
  if (!(tt > tnear && tt < tfar)) return false;
  t = tt, u = uu, v = vv;
  return true;
}

// world -> instance-local: o' = minv*(o,1), d' = minv*(d,0) in GLM's mat4*vec4 order
// (third-party/glm/glm/detail/type_mat4x4.inl:532-543): (c0*x + c1*y) + (c2*z + c3*w).  d' is NOT renormalised, so
// t is the world-space parameter (what Embree's instancing returns, EmbreeMeshAdapter.cpp:633-641).
GVT_HD void to_local(const Instance &in, V3 o, V3 d, V3 &lo, V3 &ld) {
  lo.x = xadd(xadd(xmul(in.minv0.x, o.x), xmul(in.minv0.y, o.y)), xadd(xmul(in.minv0.z, o.z), in.minv0.w));
  lo.y = xadd(xadd(xmul(in.minv1.x, o.x), xmul(in.minv1.y, o.y)), xadd(xmul(in.minv1.z, o.z), in.minv1.w));
  lo.z = xadd(xadd(xmul(in.minv2.x, o.x), xmul(in.minv2.y, o.y)), xadd(xmul(in.minv2.z, o.z), in.minv2.w));
  ld.x = xadd(xadd(xmul(in.minv0.x, d.x), xmul(in.minv0.y, d.y)), xmul(in.minv0.z, d.z));
  ld.y = xadd(xadd(xmul(in.minv1.x, d.x), xmul(in.minv1.y, d.y)), xmul(in.minv1.z, d.z));
  ld.z = xadd(xadd(xmul(in.minv2.x, d.x), xmul(in.minv2.y, d.y)), xmul(in.minv2.z, d.z));
}
// normi * v, GLM mat3*vec3 order (type_mat3x3.inl:436-439): c0[r]*x + c1[r]*y + c2[r]*z, left to right
GVT_HD V3 normi_mul(const Instance &in, V3 v) {
  return v3(xadd(xadd(xmul(in.n0.x, v.x), xmul(in.n0.y, v.y)), xmul(in.n0.z, v.z)),
            xadd(xadd(xmul(in.n1.x, v.x), xmul(in.n1.y, v.y)), xmul(in.n1.z, v.z)),
            xadd(xadd(xmul(in.n2.x, v.x), xmul(in.n2.y, v.y)), xmul(in.n2.z, v.z)));
}

// ---------------------------------------------------------------------------------------------------------
// Top-level routing: RayPacketIntersection<N>::intersect (actor/RayPacket.h:111-211) evaluated per ray inside
// BVH::intersect's traversal (data/accel/BVH.h:61-135): left child first, right pushed; at a leaf every instance
// but `from` is slab-tested with update=true and the smallest entry t (strict >) wins.  Returns the instance id
// or -1, and the entry t.  Bit-exact with the oracle: same operations, same order, same comparisons.
// ---------------------------------------------------------------------------------------------------------
GVT_HD bool slab_exact(V3 lo, V3 hi, V3 o, V3 inv, float &tnear, float &tfar) {
  float lx = xmul(xsub(lo.x, o.x), inv.x), ly = xmul(xsub(lo.y, o.y), inv.y), lz = xmul(xsub(lo.z, o.z), inv.z);
  float ux = xmul(xsub(hi.x, o.x), inv.x), uy = xmul(xsub(hi.y, o.y), inv.y), uz = xmul(xsub(hi.z, o.z), inv.z);
  float minx = fastmin(lx, ux), maxx = fastmax(lx, ux);
  float miny = fastmin(ly, uy), maxy = fastmax(ly, uy);
  float minz = fastmin(lz, uz), maxz = fastmax(lz, uz);
  tnear = fastmax(fastmax(minx, miny), minz);
  tfar = fastmin(fastmin(maxx, maxy), maxz);
  return tfar > tnear;
}

#define GVT_TLAS_STACK 64
GVT_HD int tlas_route(const Tlas &tl, V3 o, V3 d, float t_max, int from, bool enter_inside, float &t_out) {
  int next = -1;
  float t = t_max; // RayPacket.h:94  t[i] = ray.t_max
  t_out = FLT_MAX;
  if (tl.n_nodes == 0) return -1;
  V3 inv = v3(xdiv(1.f, d.x), xdiv(1.f, d.y), xdiv(1.f, d.z)); // RayPacket.h:91-93
  int stack[GVT_TLAS_STACK];
  int sp = 0;
  int cur = 0;
  for (;;) {
    float4 n0 = tl.nodes[2 * cur], n1 = tl.nodes[2 * cur + 1];
    float tn, tf;
    bool hit = slab_exact(v3(n0.x, n0.y, n0.z), v3(n1.x, n1.y, n1.z), o, inv, tn, tf) && t > tn; // update=false
    int a = (int)f2u(n0.w), b = (int)f2u(n1.w);
    if (hit) {
      if (a < 0) { // leaf
        int first = ~a;
        for (int i = first; i < first + b; ++i) {
          float4 b0 = tl.boxes[2 * i], b1 = tl.boxes[2 * i + 1];
          int id = (int)f2u(b0.w);
          if (id == from) continue; // BVH.h:112
          float tn2, tf2;
          bool ok = slab_exact(v3(b0.x, b0.y, b0.z), v3(b1.x, b1.y, b1.z), o, inv, tn2, tf2);
          bool inside = enter_inside && !(tn2 > GVT_RAY_EPSILON) && ok && tf2 > GVT_RAY_EPSILON; // deviation F6
          if (inside) tn2 = 0.f;
          ok = ok && (tn2 > GVT_RAY_EPSILON || inside) && t > tn2; // RayPacket.h:194, update=true
          if (ok) {
            t = tn2; // RayPacket.h:202
            next = id, t_out = tn2; // BVH.h:117-120 (ret.t > rp.t always holds here)
          }
        }
      } else {
        if (sp < GVT_TLAS_STACK) stack[sp++] = b; // right pushed (BVH.h:128)
        cur = a;                                  // left first
        continue;
      }
    }
    if (sp == 0) break;
    cur = stack[--sp];
  }
  return next;
}

// ---------------------------------------------------------------------------------------------------------
// 8-wide compressed BVH traversal, written as a resumable per-ray state machine so that the trace kernel can
// interleave rays (lane refill) and postpone triangle tests (Ylitie, Karras, Laine, HPG 2017, sections 4.3-4.4):
//   trav_begin   : empty state, root pending
//   node_step    : visit the nearest pending child node, test its 8 quantised child boxes
//   tri_step     : test one triangle of the current triangle group
//   trav_next    : pop the next group, or report that the traversal is finished
// Node layout: bvh_build.h.  A plane value is the float m = 1 + q/128 whose upper 16 bits are stored; t = m * A + B with
// A = 128 * step / d and B = (origin - o) / d - A.  The value of slots 0-3 is the plane word read as a float as it is
// (the lower field adds 0.2480 .. 0.2500 quanta; B is lowered by the bound that keeps near planes early and far planes
// late), the value of slots 4-7 is the word shifted left by 16 (an IMAD on the FMA pipe): no I2F, no byte extraction on the ALU pipe, which is what bounds this kernel on sm_100.
// Near / far plane blocks are picked by LOAD ADDRESS from the ray's direction signs, not by selects.  Hit slots are
// collected as an 8-bit mask: internal hits are visited in descending (slot ^ octinv) order = front to back; triangle
// bits are (leaf_hits * 0x010101) & tmask.  The box tests are free-form float math (conservative boxes), the triangle
// test is the canonical exact sequence.
// ---------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define GVT_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define GVT_LDG4(p) __ldg(p)
#else
#define GVT_FMA(a, b, c) fmaf((a), (b), (c))
#define GVT_LDG4(p) (*(p))
#endif
/* what the lower field adds to the upper value, in quanta: (0x3F80 + q') / 65536, q' in 0..127 */
#define GVT_LOWFIELD_MIN 0.24804f
#define GVT_LOWFIELD_MAX 0.24999f

#define GVT_BLAS_STACK 32
#define GVT_TRI_TAG 0x80000000u /* .x bit of a triangle group */

// the ray in mesh space, with what every node visit needs
struct RayLocal {
  V3 o, d;
  float rdx, rdy, rdz; // 1 / d, with zero components replaced by +-1e-20 (the triangle test sees the true d)
  uint32_t octinv;     // 7 - octant of d (bit 2/1/0 set = d.x/d.y/d.z >= 0)
};
GVT_HD void ray_local(RayLocal &rl, V3 o, V3 d) {
  rl.o = o, rl.d = d;
  const float tiny = 1e-20f;
  float dxs = fabsf(d.x) < tiny ? (d.x < 0.f ? -tiny : tiny) : d.x;
  float dys = fabsf(d.y) < tiny ? (d.y < 0.f ? -tiny : tiny) : d.y;
  float dzs = fabsf(d.z) < tiny ? (d.z < 0.f ? -tiny : tiny) : d.z;
  rl.rdx = 1.0f / dxs, rl.rdy = 1.0f / dys, rl.rdz = 1.0f / dzs;
  rl.octinv = 7u - ((d.x < 0.f ? 4u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 1u : 0u));
}

// the pending slot to visit first: the set bit s of hits8 with the largest (s ^ octinv)
GVT_HD uint32_t best_slot(uint32_t hits8, uint32_t octinv) {
  uint32_t best = 0, bestkey = 0;
  for (uint32_t s = 0; s < 8; ++s)
    if ((hits8 >> s) & 1u) {
      uint32_t key = (s ^ octinv) + 1u;
      if (key > bestkey) bestkey = key, best = s;
    }
  return best;
}

// traversal state of one ray in one BLAS
struct Trav {
  uint2 ng;   // node group {child base, hits8 << 24 | imask}: pending internal children of one wide node
  uint2 tg;   // triangle group {triangle base | GVT_TRI_TAG, leaf_mask << 24 | 24 triangle bits}
  int sp;     // stack entries in use
  float tmax; // far limit for the box tests (shrinks with the closest hit)
  Hit hit;
};
GVT_HD void trav_begin(Trav &tv, unsigned n_tris, float tfar) {
  tv.hit.prim = -1, tv.hit.slot = -1, tv.hit.t = tfar, tv.hit.u = 0.f, tv.hit.v = 0.f;
  tv.tmax = tfar;
  tv.sp = 0;
  // root: "child 0 of a virtual parent whose only internal slot is slot 0" -> node 0
  tv.ng = make_uint2(0u, n_tris ? 0x01000001u : 0u);
  tv.tg = make_uint2(GVT_TRI_TAG, 0u);
}
GVT_HD bool trav_has_nodes(const Trav &tv) { return (tv.ng.y >> 24) != 0u; }
GVT_HD bool trav_has_tris(const Trav &tv) { return (tv.tg.y & 0x00FFFFFFu) != 0u; }

// Visits the nearest pending child of the node group.  Precondition: trav_has_nodes(tv).  lut: optional table
// lut[octinv * 256 + hits8] = best_slot(hits8, octinv) (the trace kernel keeps it in shared memory).
// prefetch = the warp is nearly empty (the tail of a frame: what is left is a few long rays whose time is a chain of dependent
// node fetches from HBM): the siblings that go onto the stack are requested into L2 now, so that the pop that visits them
// later waits for L2 instead of HBM.  Not worth its traffic while the warp is full and the kernel is bandwidth bound.
template <bool COUNT, bool LUT, class Stack>
GVT_HD void node_step(const float4 *nodes, const RayLocal &rl, float tnear, Trav &tv, Stack &stack, TravCounters *ctr,
                      const unsigned char *lut = nullptr, bool prefetch = false) {
  const uint32_t hits8 = tv.ng.y >> 24;
  const uint32_t imask = tv.ng.y & 0xFFu;
  const uint32_t s = LUT ? (uint32_t)lut[rl.octinv * 256u + hits8] : best_slot(hits8, rl.octinv);
  tv.ng.y &= ~(0x01000000u << s);
  if (trav_has_nodes(tv)) {
    stack.push(tv.sp, tv.ng);
#if defined(__CUDA_ARCH__)
    if (prefetch) {
      uint32_t rest = tv.ng.y >> 24;
      while (rest) {
        const uint32_t b = (uint32_t)__ffs((int)rest) - 1u;
        rest &= rest - 1u;
        const float4 *pn = nodes + GVT_NODE_WORDS * (size_t)(tv.ng.x + (uint32_t)popc32(imask & ((1u << b) - 1u)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pn));
      }
    }
#endif
  }
  const uint32_t ni = tv.ng.x + (uint32_t)popc32(imask & ((1u << s) - 1u));
  const float4 *np = nodes + GVT_NODE_WORDS * (size_t)ni;
  // direction signs pick the near / far plane block of every axis
  const uint32_t sx = (rl.octinv >> 2) & 1u, sy = (rl.octinv >> 1) & 1u, sz = rl.octinv & 1u; // 1 = positive direction
  const float4 h0 = GVT_LDG4(np), h1 = GVT_LDG4(np + 1);
  const float4 nx = GVT_LDG4(np + 3 - sx), fx = GVT_LDG4(np + 2 + sx);
  const float4 ny = GVT_LDG4(np + 5 - sy), fy = GVT_LDG4(np + 4 + sy);
  const float4 nz = GVT_LDG4(np + 7 - sz), fz = GVT_LDG4(np + 6 + sz);
  if (COUNT) ctr->nodes++;
  const uint32_t ew = f2u(h0.w);
  const float ax = u2f((ew & 0xFFu) << 23) * rl.rdx, ay = u2f(((ew >> 8) & 0xFFu) << 23) * rl.rdy, az = u2f(((ew >> 16) & 0xFFu) << 23) * rl.rdz;
  const float bx = GVT_FMA(h0.x - rl.o.x, rl.rdx, -ax), by = GVT_FMA(h0.y - rl.o.y, rl.rdy, -ay), bz = GVT_FMA(h0.z - rl.o.z, rl.rdz, -az);
  // upper-half values carry +c quanta, c in [MIN, MAX]: near planes must not come out late, far planes not early.
  // t_err = (c - c_comp) * A / 128: near wants it <= 0, far >= 0
  const float kmin = -GVT_LOWFIELD_MIN / 128.f, kmax = -GVT_LOWFIELD_MAX / 128.f;
  const float bxn = GVT_FMA(ax, sx ? kmax : kmin, bx), byn = GVT_FMA(ay, sy ? kmax : kmin, by), bzn = GVT_FMA(az, sz ? kmax : kmin, bz);
  const float bxf = GVT_FMA(ax, sx ? kmin : kmax, bx), byf = GVT_FMA(ay, sy ? kmin : kmax, by), bzf = GVT_FMA(az, sz ? kmin : kmax, bz);
  const float tmax = tv.tmax;
  uint32_t hit8 = 0u;
#define GVT_LOWF(w) u2f(f2u(w) << 16)
#define GVT_PAIR(j, comp)                                                                                                \
  {                                                                                                                      \
    const float cminH = fmaxf(fmaxf(GVT_FMA(nx.comp, ax, bxn), GVT_FMA(ny.comp, ay, byn)), fmaxf(GVT_FMA(nz.comp, az, bzn), tnear)); \
    const float cmaxH = fminf(fminf(GVT_FMA(fx.comp, ax, bxf), GVT_FMA(fy.comp, ay, byf)), fminf(GVT_FMA(fz.comp, az, bzf), tmax));  \
    if (cminH <= cmaxH) hit8 |= 1u << (j);                                                                               \
    const float cminL = fmaxf(fmaxf(GVT_FMA(GVT_LOWF(nx.comp), ax, bx), GVT_FMA(GVT_LOWF(ny.comp), ay, by)),             \
                              fmaxf(GVT_FMA(GVT_LOWF(nz.comp), az, bz), tnear));                                         \
    const float cmaxL = fminf(fminf(GVT_FMA(GVT_LOWF(fx.comp), ax, bx), GVT_FMA(GVT_LOWF(fy.comp), ay, by)),             \
                              fminf(GVT_FMA(GVT_LOWF(fz.comp), az, bz), tmax));                                          \
    if (cminL <= cmaxL) hit8 |= 1u << ((j) + 4);                                                                         \
  }
  GVT_PAIR(0, x) GVT_PAIR(1, y) GVT_PAIR(2, z) GVT_PAIR(3, w)
#undef GVT_PAIR
#undef GVT_LOWF
  const uint32_t tmask = f2u(h1.z);
  tv.ng.x = f2u(h1.x);
  tv.ng.y = ((hit8 & (ew >> 24)) << 24) | (ew >> 24);
  tv.tg.x = f2u(h1.y) | GVT_TRI_TAG;
  tv.tg.y = ((tmask & 0xFFu) << 24) | ((hit8 * 0x010101u) & tmask);
}

// Tests the next triangle of the triangle group.  Precondition: trav_has_tris(tv).  Returns true when an any-hit query
// is satisfied (the traversal is over).  Bit 8k+s = k-th triangle of leaf slot s, stored at
// tri_base + 3 * (number of leaf slots below s) + k.
template <bool COUNT>
GVT_HD bool tri_step(const float4 *tris, const RayLocal &rl, float tnear, float tfar, bool any_hit, Trav &tv, TravCounters *ctr) {
  const int ti = bfind32(tv.tg.y & 0x00FFFFFFu);
  tv.tg.y &= ~(1u << ti);
  const uint32_t ls = (uint32_t)ti & 7u, lk = (uint32_t)ti >> 3;
  const uint32_t slot = (tv.tg.x & ~GVT_TRI_TAG) + 3u * (uint32_t)popc32((tv.tg.y >> 24) & ((1u << ls) - 1u)) + lk;
  const float4 *tp = tris + 3 * (size_t)slot;
  const float4 a = GVT_LDG4(tp), b = GVT_LDG4(tp + 1), c = GVT_LDG4(tp + 2);
  if (COUNT) ctr->tris++;
  float t, u, v;
  if (tri_test(rl.o, rl.d, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), tnear, tfar, t, u, v)) {
    const int prim = (int)f2u(a.w);
    // closest hit; equal t -> smaller primitive id (BVH-order independent, same rule as the oracle)
    if (any_hit || t < tv.hit.t || (t == tv.hit.t && (tv.hit.prim < 0 || prim < tv.hit.prim))) {
      tv.hit.prim = prim, tv.hit.slot = (int)slot, tv.hit.t = t, tv.hit.u = u, tv.hit.v = v;
      tv.tmax = t;
      return any_hit;
    }
  }
  return false;
}

// After the node / triangle work of one iteration: if the node group is used up, pop the next group.  A popped
// triangle group (postponed earlier) becomes the current triangle group.  Returns false when nothing is left.
template <class Stack> GVT_HD bool trav_next(Trav &tv, Stack &stack) {
  if (!trav_has_nodes(tv)) {
    if (trav_has_tris(tv)) return true; // a triangle group kept in place is still pending
    if (tv.sp == 0) return false;
    uint2 g = stack.pop(tv.sp);
    if (g.x & GVT_TRI_TAG)
      tv.tg = g, tv.ng = make_uint2(0u, 0u);
    else
      tv.ng = g;
  }
  return true;
}

// plain per-thread stack (host emulation, parity kernels)
struct LocalStack {
  uint2 e[GVT_BLAS_STACK];
  bool overflow = false;
  GVT_HD void push(int &sp, uint2 v) {
    if (sp < GVT_BLAS_STACK)
      e[sp++] = v;
    else
      overflow = true;
  }
  GVT_HD uint2 pop(int &sp) { return e[--sp]; }
};

// One ray, start to finish (parity hooks and the host emulation; the trace kernel drives the steps itself).
template <bool ANY_HIT, bool COUNT>
GVT_HD bool blas_traverse(const Blas &bl, V3 o, V3 d, float tnear, float tfar, Hit &hit, TravCounters *ctr) {
  RayLocal rl;
  ray_local(rl, o, d);
  Trav tv;
  trav_begin(tv, bl.n_tris, tfar);
  LocalStack stack;
  bool more = trav_has_nodes(tv);
  while (more) {
    if (trav_has_nodes(tv)) node_step<COUNT, false>(bl.nodes, rl, tnear, tv, stack, ctr);
    bool done = false;
    while (trav_has_tris(tv) && !done) done = tri_step<COUNT>(bl.tris, rl, tnear, tfar, ANY_HIT, tv, ctr);
    if (done) break;
    more = trav_next(tv, stack);
  }
  hit = tv.hit;
  return hit.prim >= 0;
}

} // namespace gvt
