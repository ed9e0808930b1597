// gvt_device.h -- device-side records and exact-arithmetic helpers shared by every kernel.
//
// Functions marked GVT_HD are per-thread bodies.  nvcc compiles them for sm_100a; tests/emul compiles the same
// headers with g++ (GVT_HD = inline, -ffp-contract=off) to check the BVH build / traversal LOGIC on a box
// without a GPU.  That harness lives in tests/ and is never linked into libgvtb200.so.
//
// "x*" ops (xadd, xmul, ...) are individually rounded IEEE fp32 operations that the compiler may not contract
// into FMAs: they are used wherever results must agree bit for bit with the strict-IEEE oracle (camera rays,
// routing, origin advance, Morton keys, ray/triangle test, shading).  Free-form float math (may be fused) is
// only used for the conservative box tests inside the BLAS traversal.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define GVT_HD __host__ __device__ __forceinline__
#else
#define GVT_HD inline
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float3 make_float3(float x, float y, float z) { return float3{ x, y, z }; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
static inline int2 make_int2(int x, int y) { return int2{ x, y }; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{ x, y, z, w }; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{ x, y }; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }
#endif

namespace gvt {

// ---- exact single-rounding ops -------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
GVT_HD float xadd(float a, float b) { return __fadd_rn(a, b); }
GVT_HD float xsub(float a, float b) { return __fsub_rn(a, b); }
GVT_HD float xmul(float a, float b) { return __fmul_rn(a, b); }
GVT_HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
GVT_HD float xsqrt(float a) { return __fsqrt_rn(a); }
GVT_HD float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
GVT_HD uint32_t f2u(float f) { return __float_as_uint(f); }
GVT_HD float u2f(uint32_t u) { return __uint_as_float(u); }
GVT_HD int clz32(uint32_t v) { return __clz((int)v); }
GVT_HD int clz64(uint64_t v) { return __clzll((long long)v); }
GVT_HD int popc32(uint32_t v) { return __popc(v); }
#else
// host build (emulation harness and host-side setup): compiled with -ffp-contract=off, no fast-math
GVT_HD float xadd(float a, float b) { volatile float r = a + b; return r; }
GVT_HD float xsub(float a, float b) { volatile float r = a - b; return r; }
GVT_HD float xmul(float a, float b) { volatile float r = a * b; return r; }
GVT_HD float xdiv(float a, float b) { volatile float r = a / b; return r; }
GVT_HD float xsqrt(float a) { volatile float r = sqrtf(a); return r; }
GVT_HD float xfma(float a, float b, float c) { return fmaf(a, b, c); }
GVT_HD uint32_t f2u(float f) { union { float f; uint32_t u; } c; c.f = f; return c.u; }
GVT_HD float u2f(uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }
GVT_HD int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
GVT_HD int clz64(uint64_t v) { return v ? __builtin_clzll(v) : 64; }
GVT_HD int popc32(uint32_t v) { return __builtin_popcount(v); }
#endif
GVT_HD int bfind32(uint32_t v) { return 31 - clz32(v); } // index of the highest set bit

struct V3 {
  float x, y, z;
};
GVT_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GVT_HD V3 xadd(V3 a, V3 b) { return v3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)); }
GVT_HD V3 xsub(V3 a, V3 b) { return v3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
GVT_HD V3 xmul(V3 a, float s) { return v3(xmul(a.x, s), xmul(a.y, s), xmul(a.z, s)); }
GVT_HD V3 xmul(V3 a, V3 b) { return v3(xmul(a.x, b.x), xmul(a.y, b.y), xmul(a.z, b.z)); }
GVT_HD V3 xneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
// GLM 0.9.8.1 operation order (third-party/glm/glm/detail/func_geometric.inl:54-60,76-84,94)
GVT_HD float xdot(V3 a, V3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
GVT_HD V3 xcross(V3 x, V3 y) {
  return v3(xsub(xmul(x.y, y.z), xmul(y.y, x.z)), xsub(xmul(x.z, y.x), xmul(y.z, x.x)), xsub(xmul(x.x, y.y), xmul(y.x, x.y)));
}
GVT_HD V3 xnormalize(V3 v) { return xmul(v, xdiv(1.0f, xsqrt(xdot(v, v)))); }
GVT_HD float xlength(V3 v) { return xsqrt(xdot(v, v)); }
// actor/RayPacket.h:50,58 -- NaN-order-sensitive selects, not fminf/fmaxf
GVT_HD float fastmin(float a, float b) { return (a < b) ? a : b; }
GVT_HD float fastmax(float a, float b) { return (a > b) ? a : b; }

#define GVT_RAY_EPSILON 1.e-6f /* actor/Ray.cpp:33 */
enum { RAY_PRIMARY = 0, RAY_SHADOW = 1, RAY_SECONDARY = 2,
       RAY_CAMERA_REF = 3 }; // (inbound queues of the persistent DomainScheduler only: a camera ray sent as its index, see k_persist)
enum { FLAG_SMOOTH = 1, FLAG_ENTER_INSIDE = 2, FLAG_SHADOW_TO_LIGHT = 8 };

// ---- the ray, as held in registers -----------------------------------------------------------------------
struct Ray {
  V3 o;
  float t_min;
  V3 d;
  float t_max;
  V3 color;
  float t;
  int id, depth;
  float w;
  int type;
};

// SoA ray queue in HBM: four 16-byte columns (the reference's 64-byte payload, actor/Ray.h:68-79) plus the
// target-instance column.  Every access is one coalesced 16-byte load/store per lane.
struct RayQueue {
  float4 *c0;      // origin.xyz, t_min
  float4 *c1;      // direction.xyz, t_max
  float4 *c2;      // color.rgb, t
  int4 *c3;        // id, depth, w (bits), type
  int *inst;       // instance the ray is queued for
  unsigned *count; // device counter
  unsigned *ready; // inbound queues of the persistent DomainScheduler only: slot i holds the frame number once its ray is complete
  unsigned cap;
};

GVT_HD void queue_store(const RayQueue &q, unsigned i, const Ray &r, int inst) {
  q.c0[i] = make_float4(r.o.x, r.o.y, r.o.z, r.t_min);
  q.c1[i] = make_float4(r.d.x, r.d.y, r.d.z, r.t_max);
  q.c2[i] = make_float4(r.color.x, r.color.y, r.color.z, r.t);
  q.c3[i] = make_int4(r.id, r.depth, (int)f2u(r.w), r.type);
  q.inst[i] = inst;
}
GVT_HD void queue_load(const RayQueue &q, unsigned i, Ray &r, int &inst) {
  float4 a = q.c0[i], b = q.c1[i], c = q.c2[i];
  int4 d = q.c3[i];
  r.o = v3(a.x, a.y, a.z), r.t_min = a.w;
  r.d = v3(b.x, b.y, b.z), r.t_max = b.w;
  r.color = v3(c.x, c.y, c.z), r.t = c.w;
  r.id = d.x, r.depth = d.y, r.w = u2f((uint32_t)d.z), r.type = d.w;
  inst = q.inst[i];
}

// ---- BLAS: 8-wide compressed BVH ------------------------------------------------------------------------
// One node = 80 bytes = five 16-byte words (coalesced LDG.128 each):
//   w0: origin.xyz (fp32), {ex, ey, ez, imask} bytes      ex = biased power-of-two exponent of the grid step
//   w1: child_base, tri_base, meta[0..3], meta[4..7]      imask bit s = slot s holds an internal child
//   w2: qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7]
//   w3: qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7]
//   w4: qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7]
// meta: 0 = empty; internal child in slot s: (1<<5)|(24+s); leaf child: (unary triangle count << 5) | first
// triangle offset (0..23) relative to tri_base.  Layout after Ylitie, Karras, Laine, "Efficient Incoherent Ray
// Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017); the traversal is in traverse.h.
// Triangle record = 48 bytes: {v0.xyz, prim id}, {e1.xyz, 0}, {e2.xyz, 0} with e1 = v1 - v0, e2 = v2 - v0.
struct Blas {
  const float4 *nodes; // 5 per node
  const float4 *tris;  // 3 per triangle, wide-node order
  const int4 *faces;   // per primitive (input order): I, J, K, material index (or -1)
  const float4 *normals; // per vertex (xyz, 0) or null
  const float4 *vcolors; // per vertex or null
  const float4 *mats;    // 2 per material: {type, kd.xyz}, {ks.xyz, alpha}
  float4 mat0, mat1;     // mesh material
  unsigned n_tris;
  unsigned n_nodes;
};

struct Hit {
  int prim;  // primitive id (input order), -1 = miss
  int slot;  // index of the triangle record
  float t, u, v;
};

struct TravCounters {
  unsigned nodes, tris;
};

// ---- scene ----------------------------------------------------------------------------------------------
struct Instance {
  float4 minv0, minv1, minv2; // rows of the 3x4 inverse matrix: row r = (c0[r], c1[r], c2[r], c3[r])
  float4 n0, n1, n2;          // rows of normi (mat3): row r = (c0[r], c1[r], c2[r], 0)
  int mesh;                   // index into the Blas table, -1 if not resident on this GPU
  int owner;                  // DomainScheduler: owning rank
  int pad0, pad1;
};

// top-level BVH over instance boxes, flattened from data/accel/BVH.cpp's tree.  32-byte nodes:
//   lo.xyz, a      a >= 0: left child index;  a < 0: leaf, first sorted instance = ~a
//   hi.xyz, b      internal: right child index; leaf: number of instances
struct Tlas {
  const float4 *nodes;  // 2 per node
  const float4 *boxes;  // 2 per sorted instance: {lo.xyz, id}, {hi.xyz, 0}
  int n_nodes;
  int n_inst;
};

struct Light {
  int type;
  V3 pos, color, normal;
  float width, height;
};
#define GVT_MAX_LIGHTS 8

struct Camera {
  V3 eye, u, v, w;
  float horz, vert, wmult, hmult, half_sample, offset, contri;
  int W, H, samples, depth;
};

struct FrameParams {
  Camera cam;
  Light lights[GVT_MAX_LIGHTS];
  int n_lights;
  unsigned flags, seed;
  int rank, world; // DomainScheduler; world == 1 => everything is local
};

struct DevCounters {
  unsigned long long camera_rays, routed_initial, primary_hits, shadow_generated, shadow_unoccluded, secondary_generated,
      fb_adds, instance_visits, nodes_visited, tris_tested, rays_exported, stack_overflows, queue_overflows;
};

// ---- counter-based RNG (replaces the per-TBB-chunk stateful generator) ------------------------
GVT_HD uint32_t pcg(uint32_t v) {
  uint32_t s = v * 747796405u + 2891336453u;
  uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
  return (w >> 22u) ^ w;
}
GVT_HD float rnd01(uint32_t seed, int id, int depth, int dim) {
  uint32_t h = pcg(seed ^ pcg((uint32_t)id * 4u + 0x9e3779b9u * (uint32_t)(depth + 1)) ^ (0x85ebca6bu * (uint32_t)(dim + 1)));
  return xmul((float)(h >> 8), (1.0f / 16777216.0f));
}

} // namespace gvt
