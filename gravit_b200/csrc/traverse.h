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
// 8-wide compressed BVH traversal.  One thread per ray; a small stack of (node group | triangle group) pairs.
// Child boxes are dequantised in the ray's frame: t = q * (2^e / d) + (p - o) / d.  The box tests are free-form
// float math (conservative boxes: see bvh_build.h), the triangle test is the canonical exact sequence.
// ---------------------------------------------------------------------------------------------------------
GVT_HD uint32_t sign_extend_s8x4(uint32_t x) { return ((x >> 7) & 0x01010101u) * 0xFFu; }
GVT_HD uint32_t byte_of(uint32_t w, int j) { return (w >> (8 * j)) & 0xFFu; }

#define GVT_BLAS_STACK 32

template <bool ANY_HIT, bool COUNT>
GVT_HD bool blas_traverse(const Blas &bl, V3 o, V3 d, float tnear, float tfar, Hit &hit, TravCounters *ctr) {
  hit.prim = -1, hit.slot = -1, hit.t = tfar, hit.u = 0.f, hit.v = 0.f;
  if (bl.n_tris == 0) return false;
  // slab directions: avoid 1/0 (the triangle test still sees the true direction)
  const float tiny = 1e-20f;
  float dxs = fabsf(d.x) < tiny ? (d.x < 0.f ? -tiny : tiny) : d.x;
  float dys = fabsf(d.y) < tiny ? (d.y < 0.f ? -tiny : tiny) : d.y;
  float dzs = fabsf(d.z) < tiny ? (d.z < 0.f ? -tiny : tiny) : d.z;
  const float rdx = 1.0f / dxs, rdy = 1.0f / dys, rdz = 1.0f / dzs;
  const uint32_t octinv = 7u - ((d.x < 0.f ? 4u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 1u : 0u));
  const uint32_t octinv4 = octinv * 0x01010101u;
  float tmax = tfar;
  uint2 stack[GVT_BLAS_STACK];
  int sp = 0;
  uint2 ngroup = make_uint2(0u, 0x80000000u); // root: node 0, one pending child bit
  uint2 tgroup = make_uint2(0u, 0u);
  for (;;) {
    { // invariant: ngroup.y > 0x00FFFFFF here (a node group with at least one pending child)
      const uint32_t hits = ngroup.y;
      const uint32_t imask = ngroup.y; // low byte
      const int child_bit_index = bfind32(hits);
      const uint32_t child_node_base = ngroup.x;
      ngroup.y &= ~(1u << child_bit_index);
      if (ngroup.y > 0x00FFFFFFu) {
        if (sp < GVT_BLAS_STACK) stack[sp++] = ngroup;
      }
      const uint32_t slot_index = ((uint32_t)child_bit_index - 24u) ^ (octinv4 & 255u);
      const uint32_t relative = (uint32_t)popc32(imask & ~(0xFFFFFFFFu << slot_index));
      const uint32_t ni = child_node_base + relative;
      const float4 n0 = bl.nodes[5 * ni + 0], n1 = bl.nodes[5 * ni + 1], n2 = bl.nodes[5 * ni + 2],
                   n3 = bl.nodes[5 * ni + 3], n4 = bl.nodes[5 * ni + 4];
      if (COUNT) ctr->nodes++;
      const uint32_t ew = f2u(n0.w);
      ngroup.x = f2u(n1.x);
      tgroup.x = f2u(n1.y);
      tgroup.y = 0u;
      uint32_t hitmask = 0u;
      const float adjx = u2f((ew & 0xFFu) << 23) * rdx;
      const float adjy = u2f(((ew >> 8) & 0xFFu) << 23) * rdy;
      const float adjz = u2f(((ew >> 16) & 0xFFu) << 23) * rdz;
      const float ox = (n0.x - o.x) * rdx, oy = (n0.y - o.y) * rdy, oz = (n0.z - o.z) * rdz;
      const uint32_t qw[12] = { f2u(n2.x), f2u(n2.y), f2u(n2.z), f2u(n2.w), f2u(n3.x), f2u(n3.y),
                                f2u(n3.z), f2u(n3.w), f2u(n4.x), f2u(n4.y), f2u(n4.z), f2u(n4.w) };
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? f2u(n1.w) : f2u(n1.z);
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        // words: qlo.x = qw[0+half], qlo.y = qw[2+half], qlo.z = qw[4+half], qhi.x = qw[6+half], qhi.y = qw[8+half], qhi.z = qw[10+half]
        const uint32_t lox = rdx < 0.f ? qw[6 + half] : qw[0 + half], hix = rdx < 0.f ? qw[0 + half] : qw[6 + half];
        const uint32_t loy = rdy < 0.f ? qw[8 + half] : qw[2 + half], hiy = rdy < 0.f ? qw[2 + half] : qw[8 + half];
        const uint32_t loz = rdz < 0.f ? qw[10 + half] : qw[4 + half], hiz = rdz < 0.f ? qw[4 + half] : qw[10 + half];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) {
          const float tlx = (float)byte_of(lox, j) * adjx + ox, tly = (float)byte_of(loy, j) * adjy + oy,
                      tlz = (float)byte_of(loz, j) * adjz + oz;
          const float thx = (float)byte_of(hix, j) * adjx + ox, thy = (float)byte_of(hiy, j) * adjy + oy,
                      thz = (float)byte_of(hiz, j) * adjz + oz;
          const float cmin = fmaxf(fmaxf(tlx, tly), fmaxf(tlz, tnear));
          const float cmax = fminf(fminf(thx, thy), fminf(thz, tmax));
          if (cmin <= cmax) hitmask |= byte_of(child_bits4, j) << byte_of(bit_index4, j);
        }
      }
      ngroup.y = (hitmask & 0xFF000000u) | ((ew >> 24) & 0xFFu);
      tgroup.y = hitmask & 0x00FFFFFFu;
    }
    while (tgroup.y != 0u) {
      const int ti = bfind32(tgroup.y);
      tgroup.y &= ~(1u << ti);
      const uint32_t slot = tgroup.x + (uint32_t)ti;
      const float4 a = bl.tris[3 * slot], b = bl.tris[3 * slot + 1], c = bl.tris[3 * slot + 2];
      if (COUNT) ctr->tris++;
      float t, u, v;
      if (tri_test(o, d, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), tnear, tfar, t, u, v)) {
        const int prim = (int)f2u(a.w);
        if (ANY_HIT) {
          hit.prim = prim, hit.slot = (int)slot, hit.t = t, hit.u = u, hit.v = v;
          return true;
        }
        // closest hit; equal t -> smaller primitive id (BVH-order independent, same rule as the oracle)
        if (t < hit.t || (t == hit.t && (hit.prim < 0 || prim < hit.prim))) {
          hit.prim = prim, hit.slot = (int)slot, hit.t = t, hit.u = u, hit.v = v;
          tmax = t;
        }
      }
    }
    if (ngroup.y <= 0x00FFFFFFu) {
      if (sp == 0) break;
      ngroup = stack[--sp];
    }
  }
  return hit.prim >= 0;
}

} // namespace gvt
