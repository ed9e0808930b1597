// shade.h -- per-thread bodies for camera-ray generation, hit shading, shadow-ray and bounce generation.
// All arithmetic here is the exact single-rounding sequence of the oracle (oracle/gvt_oracle.cpp), which in turn
// follows GLM's operation order, so rays and colours agree bit for bit (Lambert; Phong/Blinn use powf and agree
// to a few ulp).
#pragma once
#include "gvt_device.h"
#include "traverse.h"

namespace gvt {

// gvtPerspectiveCamera::generateRays (data/scene/gvtCamera.cpp:233-312); ridx = (j*W+i)*samples^2 + k*samples + w
GVT_HD void camera_ray(const Camera &c, size_t ridx, Ray &ray) {
  const int s2 = c.samples * c.samples;
  int idx = (int)(ridx / (size_t)s2);
  int sub = (int)(ridx - (size_t)idx * s2);
  int k = sub / c.samples, ww = sub - k * c.samples;
  int j = idx / c.W, i = idx - j * c.W;
  float x0 = xsub(xmul((float)i, c.wmult), 1.0f), y0 = xsub(xmul((float)j, c.hmult), 1.0f);
  float x = xadd(x0, xmul(xsub((float)ww, c.half_sample), c.offset));
  x = xmul(x, c.horz);
  float y = xadd(y0, xmul(xsub((float)k, c.half_sample), c.offset));
  y = xmul(y, c.vert);
  V3 dir;
  dir.x = xadd(xadd(xmul(c.u.x, x), xmul(c.v.x, y)), c.w.x);
  dir.y = xadd(xadd(xmul(c.u.y, x), xmul(c.v.y, y)), c.w.y);
  dir.z = xadd(xadd(xmul(c.u.z, x), xmul(c.v.z, y)), c.w.z);
  ray.id = idx;
  ray.t_min = GVT_RAY_EPSILON;
  ray.o = c.eye;
  ray.d = xnormalize(dir);
  ray.t_max = FLT_MAX;
  ray.color = v3(0.f, 0.f, 0.f);
  ray.t = FLT_MAX;
  ray.w = c.contri;
  ray.type = RAY_PRIMARY;
  ray.depth = c.depth;
}

struct Mat {
  int type;
  V3 kd, ks;
  float alpha;
};
GVT_HD Mat unpack_mat(float4 a, float4 b) {
  Mat m;
  m.type = (int)f2u(a.x);
  m.kd = v3(a.y, a.z, a.w);
  m.ks = v3(b.x, b.y, b.z);
  m.alpha = b.w;
  return m;
}

GVT_HD float stdmax(float a, float b) { return (a < b) ? b : a; } // std::max
GVT_HD float clamp01(float x) {
  float a = (x < 0.f) ? 0.f : x;
  return (1.f < a) ? 1.f : a;
}

// Light::contribution (data/scene/Light.cpp:58-62,72,129-133)
GVT_HD V3 light_contribution(const Light &L, V3 hit, V3 samplePos) {
  if (L.type == 1) return L.color;
  V3 p = (L.type == 2) ? samplePos : L.pos;
  float distance = xdiv(1.f, xlength(xsub(p, hit)));
  distance = (distance > 1.f) ? 1.f : distance;
  return xmul(L.color, distance);
}

// Shade (data/primitives/Material.cpp:90-139) with lambert/phong/blinn (:50-86)
GVT_HD bool shade(const Mat &mat, const Ray &ray, V3 N, const Light &L, V3 lightPos, V3 &color) {
  V3 hitPoint = xadd(ray.o, xmul(ray.d, ray.t));
  V3 wi = xnormalize(xsub(lightPos, hitPoint));
  float NdotL = stdmax(0.f, xdot(N, wi));
  V3 Li = light_contribution(L, hitPoint, lightPos);
  if (NdotL == 0.f || (Li.x == 0.f && Li.y == 0.f && Li.z == 0.f)) return false;
  if (mat.type == 0) {
    color = xmul(mat.kd, xmul(NdotL, ray.w));
  } else if (mat.type == 1) {
    V3 R = xsub(xmul(xmul(N, 2.f), NdotL), wi);
    float VdotR = stdmax(0.f, xdot(R, xneg(ray.d)));
    float power = xmul(VdotR, powf(VdotR, mat.alpha));
    color = xmul(mat.kd, xmul(NdotL, ray.w));
    color = xadd(color, xmul(mat.ks, xmul(power, ray.w)));
  } else if (mat.type == 2) {
    V3 H = xnormalize(xsub(wi, ray.d));
    float NdotH = stdmax(0.f, xdot(H, N));
    float power = xmul(NdotH, powf(NdotH, mat.alpha));
    V3 diffuse = xmul(mat.kd, xmul(NdotL, ray.w));
    V3 specular = xmul(mat.ks, xmul(power, ray.w));
    color = xadd(diffuse, specular);
  } else {
    color = v3(0.f, 0.f, 0.f);
  }
  color = xmul(color, Li);
  color = v3(clamp01(color.x), clamp01(color.y), clamp01(color.z));
  return true;
}

// AreaLight::GetPosition (data/scene/Light.cpp:115-127) with the counter-based RNG
GVT_HD V3 light_position(const Light &L, unsigned seed, int id, int depth, int li) {
  if (L.type != 2) return L.pos;
  float xl = xmul(xsub(rnd01(seed, id, depth, 8 + 2 * li), 0.5f), L.width);
  float zl = xmul(xsub(rnd01(seed, id, depth, 9 + 2 * li), 0.5f), L.height);
  V3 vv = L.normal, uu, ww;
  if (vv.x == 0.f && vv.y == 1.f && vv.z == 0.f) {
    uu = v3(1.f, 0.f, 0.f), ww = v3(0.f, 0.f, 1.f);
  } else { // up = (0,1,0): u = up x v, w = v x u written out as in Light.cpp:93-101
    uu = v3(xsub(xmul(1.f, vv.z), xmul(vv.y, 0.f)), xsub(xmul(0.f, vv.x), xmul(vv.z, 0.f)), xsub(xmul(0.f, vv.y), xmul(vv.x, 1.f)));
    ww = v3(xsub(xmul(vv.y, uu.z), xmul(uu.y, vv.z)), xsub(xmul(vv.z, uu.x), xmul(uu.z, vv.x)), xsub(xmul(vv.x, uu.y), xmul(uu.x, vv.y)));
  }
  return v3(xadd(L.pos.x, xadd(xmul(xl, uu.x), xmul(zl, ww.x))), xadd(L.pos.y, xadd(xmul(xl, uu.y), xmul(zl, ww.y))),
            xadd(L.pos.z, xadd(xmul(xl, uu.z), xmul(zl, ww.z))));
}

// Surface data at a hit: shading normal (flat at HEAD, smooth for the goldens) flipped towards the viewer, and
// the material (vertex colour > face material > mesh material).  EmbreeMeshAdapter.cpp:491-569.
GVT_HD void surface_at(const Blas &bl, const Instance &in, const Hit &h, const Ray &r, unsigned flags, V3 &normal, Mat &mat) {
  const float4 a = bl.tris[3 * h.slot + 1], b = bl.tris[3 * h.slot + 2];
  V3 e1 = v3(a.x, a.y, a.z), e2 = v3(b.x, b.y, b.z);
  // :504 normalflat = normalize(normi * -Ng), -Ng_embree = cross(v1-v0, v2-v0) (sign pinned by the goldens, F5)
  V3 normalflat = xnormalize(normi_mul(in, xcross(e1, e2)));
  const int4 f = bl.faces[h.prim];
  if ((flags & FLAG_SMOOTH) && bl.normals) { // :507-519  a=N[J], b=N[K], c=N[I]
    float4 nj = bl.normals[f.y], nk = bl.normals[f.z], ni = bl.normals[f.x];
    V3 m = xadd(xadd(xmul(v3(nj.x, nj.y, nj.z), h.u), xmul(v3(nk.x, nk.y, nk.z), h.v)),
                xmul(v3(ni.x, ni.y, ni.z), xsub(xsub(1.0f, h.u), h.v)));
    normal = xnormalize(normi_mul(in, m));
  } else {
    normal = normalflat;
  }
  if (xdot(xneg(r.d), normalflat) <= 0.f) normal = xneg(normal); // :528-530
  if (bl.vcolors) { // :536-563
    float4 c0 = bl.vcolors[f.x], c1 = bl.vcolors[f.y], c2 = bl.vcolors[f.z];
    mat.type = 0;
    mat.kd = xadd(xadd(xmul(v3(c0.x, c0.y, c0.z), xsub(xsub(1.f, h.u), h.v)), xmul(v3(c1.x, c1.y, c1.z), h.u)),
                  xmul(v3(c2.x, c2.y, c2.z), h.v));
    mat.ks = v3(.5f, .5f, .5f);
    mat.alpha = 1.f;
  } else if (bl.mats && f.w >= 0) { // :565-566
    mat = unpack_mat(bl.mats[2 * f.w], bl.mats[2 * f.w + 1]);
  } else {
    mat = unpack_mat(bl.mat0, bl.mat1); // :568
  }
}

// generateShadowRays for one light (EmbreeMeshAdapter.cpp:320-358).  false = Shade() rejected the light.
GVT_HD bool make_shadow_ray(const Ray &r, V3 normal, const Mat &mat, const Light &L, V3 lightPos, bool to_light, Ray &sh) {
  V3 c;
  if (!shade(mat, r, normal, L, lightPos, c)) return false;
  const float multiplier = xsub(1.0f, xmul(GVT_RAY_EPSILON, 16.f));
  const float t_shadow = xmul(multiplier, r.t);
  V3 origin = xadd(r.o, xmul(r.d, t_shadow));
  V3 dir = xsub(lightPos, origin);
  sh.o = origin;
  sh.d = xnormalize(dir); // Ray ctor, actor/Ray.h:109
  sh.t_min = GVT_RAY_EPSILON;
  sh.t_max = to_light ? xlength(dir) : 3.0f; // glm::vec3::length() == 3 (:347,355); FLAG_SHADOW_TO_LIGHT: the distance
  sh.t = r.t;
  sh.id = r.id;
  sh.w = r.w;
  sh.type = RAY_SHADOW;
  sh.depth = r.depth;
  sh.color = c;
  return true;
}

// CosWeightedRandomHemisphereDirection2 (EmbreeMeshAdapter.cpp:289-318), fp32 (PSNR-level parity)
GVT_HD V3 cos_hemisphere(V3 n, float Xi1, float Xi2) {
  float ys = sqrtf(fmaxf(0.f, 1.0f - Xi1)); // cos(acos(sqrt(1-Xi1)))
  float st = sqrtf(fmaxf(0.f, Xi1));        // sin(theta)
  float phi = 6.283185307179586f * Xi2;
  float xs = st * cosf(phi), zs = st * sinf(phi);
  V3 y = n, h = y;
  if (fabsf(h.x) <= fabsf(h.y) && fabsf(h.x) <= fabsf(h.z))
    h.x = 1.0f;
  else if (fabsf(h.y) <= fabsf(h.x) && fabsf(h.y) <= fabsf(h.z))
    h.y = 1.0f;
  else
    h.z = 1.0f;
  V3 x = xcross(h, y), z = xcross(x, y);
  V3 dir = xadd(xadd(xmul(x, xs), xmul(y, ys)), xmul(z, zs));
  return xnormalize(dir);
}

// the bounce decision of EmbreeMeshAdapter.cpp:584-601: true = r became a SECONDARY ray in place
GVT_HD bool make_bounce(Ray &r, V3 normal, unsigned seed) {
  int ndepth = r.depth - 1;
  if (ndepth <= 0) return false;
  float p = xsub(1.f, rnd01(seed, r.id, r.depth, 0));
  if (!(r.w > p)) return false;
  r.type = RAY_SECONDARY;
  const float multiplier = xsub(1.0f, xmul(16.0f, FLT_EPSILON));
  const float t_secondary = xmul(multiplier, r.t);
  r.o = xadd(r.o, xmul(r.d, t_secondary));
  r.d = cos_hemisphere(normal, rnd01(seed, r.id, r.depth, 1), rnd01(seed, r.id, r.depth, 2));
  r.w = xmul(r.w, xdot(r.d, normal));
  r.depth = ndepth;
  return true;
}

} // namespace gvt
