// api.cpp -- host-side mirror of gvt::render::api (reference: src/gvt/render/api/api.cpp:76-540) and of the pieces
// of the data model that feed the hot path: Mesh (data/primitives/Mesh.cpp:62-199), addInstance's matrices
// (api.cpp:292-322, GLM 0.9.8.1 operation order), gvtRenderer::render (Renderer.cpp:37-115) and
// IceTComposite::write (composite/IceTComposite.cpp:119-158).  All ray work is done by libgvtb200.so (CUDA, sm_100a);
// if that library cannot find a device every render throws -- there is no CPU path.
//
// Built with -ffp-contract=off: every expression below is evaluated as written, one IEEE rounding per operation.
#include "api.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <condition_variable>
#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "../../include/gvtb200.h"

namespace {

struct V3 {
  float x, y, z;
};
inline V3 operator-(V3 a, V3 b) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator+(V3 a, V3 b) { return V3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// glm::normalize = v * inversesqrt(dot(v,v)); dot = (x*x + y*y) + z*z (func_geometric.inl:54-60,94)
inline V3 normalize(V3 v) {
  float d = v.x * v.x + v.y * v.y + v.z * v.z;
  float s = 1.0f / std::sqrt(d);
  return V3{ v.x * s, v.y * s, v.z * s };
}
inline float fastmin(float a, float b) { return a < b ? a : b; } // actor/RayPacket.h:50,58 (used by Box3D)
inline float fastmax(float a, float b) { return a > b ? a : b; }

struct Material { // data/primitives/Material.h:59-89
  int type = 0;   // LAMBERT
  float kd[3] = { .5f, .5f, .5f }, ks[3] = { .5f, .5f, .5f };
  float alpha = 1.f;
};

struct Mesh { // data/primitives/Mesh.h
  std::vector<V3> vertices, normals, face_normals, vertex_colors;
  std::vector<unsigned> faces; // 0-based, 3 per face
  std::vector<Material> faces_to_materials;
  Material mat;
  bool haveNormals = false;
  float bmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, bmax[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  gvtb200_mesh device = nullptr; // the adapter object, one per mesh (TracerBase.h:449-450 adapterCache)
  bool dirty = true;

  ~Mesh() {
    if (device) gvtb200_mesh_destroy(device);
  }
  void addFace(int v0, int v1, int v2) { // Mesh.cpp:102-112: 1-based in, degenerate faces dropped
    const int nv = (int)vertices.size();
    if (!(v0 - 1 >= 0 && v0 - 1 < nv)) throw std::runtime_error("Vertex index 0 outside bounds : " + std::to_string(v0 - 1));
    if (!(v1 - 1 >= 0 && v1 - 1 < nv)) throw std::runtime_error("Vertex index 1 outside bounds : " + std::to_string(v1 - 1));
    if (!(v2 - 1 >= 0 && v2 - 1 < nv)) throw std::runtime_error("Vertex index 2 outside bounds : " + std::to_string(v2 - 1));
    if (vertices[v0 - 1] == vertices[v1 - 1] || vertices[v1 - 1] == vertices[v2 - 1] || vertices[v2 - 1] == vertices[v0 - 1]) return;
    faces.push_back(v0 - 1), faces.push_back(v1 - 1), faces.push_back(v2 - 1);
  }
  void generateNormals() { // Mesh.cpp:116-154
    if (haveNormals) return;
    normals.assign(vertices.size(), V3{ 0.f, 0.f, 0.f });
    face_normals.resize(faces.size() / 3);
    for (size_t i = 0; i < faces.size() / 3; ++i) {
      unsigned I = faces[3 * i], J = faces[3 * i + 1], K = faces[3 * i + 2];
      V3 u = vertices[J] - vertices[I], v = vertices[K] - vertices[I], n;
      n.x = u.y * v.z - u.z * v.y;
      n.y = u.z * v.x - u.x * v.z;
      n.z = u.x * v.y - u.y * v.x;
      n = normalize(n);
      face_normals[i] = n;
      normals[I] = normals[I] + n;
      normals[J] = normals[J] + n;
      normals[K] = normals[K] + n;
    }
    for (auto &n : normals) n = normalize(n);
    haveNormals = true;
  }
  void computeBoundingBox() { // Mesh.cpp:186-199 + Box3D::expand BBox.cpp:65-73
    for (int i = 0; i < 3; ++i) bmin[i] = FLT_MAX, bmax[i] = -FLT_MAX;
    for (const V3 &p : vertices) {
      bmin[0] = fastmin(bmin[0], p.x), bmin[1] = fastmin(bmin[1], p.y), bmin[2] = fastmin(bmin[2], p.z);
      bmax[0] = fastmax(bmax[0], p.x), bmax[1] = fastmax(bmax[1], p.y), bmax[2] = fastmax(bmax[2], p.z);
    }
  }
};

struct Instance {
  std::string name, mesh;
  float m[16], minv[16], normi[9], lo[3], hi[3];
};
struct LightRec {
  std::string name;
  gvtb200_light l;
};
struct CameraRec {
  float eye[3], focus[3], up[3], fov;
  int depth, samples;
  float jitter;
};
struct FilmRec {
  int w, h;
  std::string path;
};
struct RendererRec {
  int adapter, schedule;
  std::string camera, film;
  bool volume;
};

struct Context {
  std::map<std::string, std::shared_ptr<Mesh> > meshes;
  std::vector<Instance> instances; // creation order == the ids BVH::BVH assigns (data/accel/BVH.cpp:48-51)
  std::vector<LightRec> lights;    // creation order
  std::map<std::string, CameraRec> cameras;
  std::map<std::string, FilmRec> films;
  std::map<std::string, RendererRec> renderers;
  std::map<std::string, int> names; // unique names across the database (core/cntx/context.h:93-117)
  unsigned flags = 0, seed = 0;
  bool initialised = false;
  // device scene
  gvtb200_scene scene = nullptr;
  bool instances_dirty = true;
  // several GPUs in this process (api::setGpus): per GPU its own BLAS handles and scene; the tracers of algorithm/*.h run
  // as one host thread per GPU (render_multi below)
  int gpus = 1;
  struct Multi {
    int n = 0, schedule = -1, w = 0, h = 0;
    bool attached = false;
    std::vector<std::vector<gvtb200_mesh> > meshes; // [gpu][mesh in map order], nullptr = not resident
    std::vector<gvtb200_scene> scenes;
  } multi;
  void drop_multi() {
    for (size_t g = 0; g < multi.scenes.size(); ++g)
      if (multi.scenes[g]) gvtb200_scene_destroy(multi.scenes[g]);
    for (auto &v : multi.meshes)
      for (auto m : v)
        if (m) gvtb200_mesh_destroy(m);
    multi = Multi();
  }
  std::vector<float> fb;
  int fbw = 0, fbh = 0;

  void unique(const std::string &n) {
    if (names.count(n)) throw std::runtime_error("Name " + n + " is not unique"); // context.h:98
    names[n] = 1;
  }
  void drop_scene() {
    if (scene) gvtb200_scene_destroy(scene);
    scene = nullptr;
    instances_dirty = true;
    drop_multi();
  }
  ~Context() { drop_scene(); }
};
Context &db() {
  static Context c;
  return c;
}
std::shared_ptr<Mesh> mesh_of(const std::string &name) {
  auto it = db().meshes.find(name);
  if (it == db().meshes.end()) throw std::runtime_error("Node " + name + " not found"); // context.h:89
  return it->second;
}
[[noreturn]] void engine_error(const char *what) { throw std::runtime_error(std::string(what) + ": " + gvtb200_last_error()); }

// ---- GLM 0.9.8.1 matrix routines in GLM's operation order (third-party/glm/glm/detail) ---------------------
// column-major a[4*c + r]
void mat4_inverse(const float *a, float *out) { // func_matrix.inl:297-354
  auto m = [&](int c, int r) { return a[4 * c + r]; };
  float C00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), C02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), C03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
  float C04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), C06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), C07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
  float C08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), C10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), C11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
  float C12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), C14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), C15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
  float C16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), C18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), C19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
  float C20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1), C22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), C23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
  const float F0[4] = { C00, C00, C02, C03 }, F1[4] = { C04, C04, C06, C07 }, F2[4] = { C08, C08, C10, C11 };
  const float F3[4] = { C12, C12, C14, C15 }, F4[4] = { C16, C16, C18, C19 }, F5[4] = { C20, C20, C22, C23 };
  const float V0[4] = { m(1, 0), m(0, 0), m(0, 0), m(0, 0) }, V1[4] = { m(1, 1), m(0, 1), m(0, 1), m(0, 1) };
  const float V2[4] = { m(1, 2), m(0, 2), m(0, 2), m(0, 2) }, V3_[4] = { m(1, 3), m(0, 3), m(0, 3), m(0, 3) };
  const float SA[4] = { +1, -1, +1, -1 }, SB[4] = { -1, +1, -1, +1 };
  float inv[16];
  for (int i = 0; i < 4; ++i) {
    float i0 = V1[i] * F0[i] - V2[i] * F1[i] + V3_[i] * F2[i];
    float i1 = V0[i] * F0[i] - V2[i] * F3[i] + V3_[i] * F4[i];
    float i2 = V0[i] * F1[i] - V1[i] * F3[i] + V3_[i] * F5[i];
    float i3 = V0[i] * F2[i] - V1[i] * F4[i] + V2[i] * F5[i];
    inv[0 + i] = i0 * SA[i], inv[4 + i] = i1 * SB[i], inv[8 + i] = i2 * SA[i], inv[12 + i] = i3 * SB[i];
  }
  float d0 = m(0, 0) * inv[0], d1 = m(0, 1) * inv[4], d2 = m(0, 2) * inv[8], d3 = m(0, 3) * inv[12];
  float dot1 = (d0 + d1) + (d2 + d3);
  float ood = 1.0f / dot1;
  for (int i = 0; i < 16; ++i) out[i] = inv[i] * ood;
}
// transpose(inverse(mat3(m))) (func_matrix.inl:272-293), column-major 3x3 out
void mat3_inverse_transpose(const float *a, float *out) {
  auto m = [&](int c, int r) { return a[4 * c + r]; };
  float ood = 1.0f / (+m(0, 0) * (m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2)) - m(1, 0) * (m(0, 1) * m(2, 2) - m(2, 1) * m(0, 2)) +
                      m(2, 0) * (m(0, 1) * m(1, 2) - m(1, 1) * m(0, 2)));
  float I[3][3]; // I[c][r]
  I[0][0] = +(m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2)) * ood;
  I[1][0] = -(m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2)) * ood;
  I[2][0] = +(m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1)) * ood;
  I[0][1] = -(m(0, 1) * m(2, 2) - m(2, 1) * m(0, 2)) * ood;
  I[1][1] = +(m(0, 0) * m(2, 2) - m(2, 0) * m(0, 2)) * ood;
  I[2][1] = -(m(0, 0) * m(2, 1) - m(2, 0) * m(0, 1)) * ood;
  I[0][2] = +(m(0, 1) * m(1, 2) - m(1, 1) * m(0, 2)) * ood;
  I[1][2] = -(m(0, 0) * m(1, 2) - m(1, 0) * m(0, 2)) * ood;
  I[2][2] = +(m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1)) * ood;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) out[3 * c + r] = I[r][c];
}
// vec3(m * vec4(p, 1)) (type_mat4x4.inl:532-543): (c0*x + c1*y) + (c2*z + c3*w)
void mat4_point(const float *a, const float *p, float *out) {
  for (int r = 0; r < 3; ++r) {
    float mul0 = a[0 + r] * p[0], mul1 = a[4 + r] * p[1], add0 = mul0 + mul1;
    float mul2 = a[8 + r] * p[2], mul3 = a[12 + r] * 1.f, add1 = mul2 + mul3;
    out[r] = add0 + add1;
  }
}

void ensure_init() {
  if (db().initialised) return;
  if (gvtb200_current_device() < 0 && gvtb200_init(0) != GVTB200_OK) engine_error("gvtInit"); // keep a device the caller chose
  db().initialised = true;
}

// builds / refreshes the device scene from the database (what AbstractTrace::Initialize does, TracerBase.h:247-308)
gvtb200_scene device_scene(const RendererRec &r) {
  Context &c = db();
  ensure_init();
  auto cam = c.cameras.find(r.camera);
  auto film = c.films.find(r.film);
  if (cam == c.cameras.end()) throw std::runtime_error("Node " + r.camera + " not found");
  if (film == c.films.end()) throw std::runtime_error("Node " + r.film + " not found");
  for (auto &kv : c.meshes) { // one adapter (BLAS) per mesh, built on first use and cached (ImageTracer.h:177-233)
    Mesh &m = *kv.second;
    if (m.device && !m.dirty) continue;
    if (m.device) gvtb200_mesh_destroy(m.device), m.device = nullptr;
    m.generateNormals(); // EmbreeMeshAdapter::EmbreeMeshAdapter, EmbreeMeshAdapter.cpp:127
    gvtb200_material mm;
    mm.type = m.mat.type, memcpy(mm.kd, m.mat.kd, 12), memcpy(mm.ks, m.mat.ks, 12), mm.alpha = m.mat.alpha;
    std::vector<gvtb200_material> fm;
    std::vector<int32_t> fidx;
    const size_t nf = m.faces.size() / 3;
    if (!m.faces_to_materials.empty()) { // EmbreeMeshAdapter.cpp:565-566
      fm.resize(m.faces_to_materials.size());
      for (size_t i = 0; i < fm.size(); ++i) {
        const Material &s = m.faces_to_materials[i];
        fm[i].type = s.type, memcpy(fm[i].kd, s.kd, 12), memcpy(fm[i].ks, s.ks, 12), fm[i].alpha = s.alpha;
      }
      fidx.resize(nf);
      for (size_t i = 0; i < nf; ++i) fidx[i] = i < fm.size() ? (int32_t)i : -1;
    }
    const bool vc = !m.vertex_colors.empty() && m.vertex_colors.size() == m.vertices.size();
    if (gvtb200_mesh_create(&m.vertices[0].x, (uint32_t)m.vertices.size(), m.faces.data(), (uint32_t)nf,
                            m.normals.size() == m.vertices.size() ? &m.normals[0].x : nullptr, &mm, fm.empty() ? nullptr : fm.data(),
                            (uint32_t)fm.size(), fidx.empty() ? nullptr : fidx.data(), vc ? &m.vertex_colors[0].x : nullptr,
                            &m.device) != GVTB200_OK)
      engine_error("mesh build");
    m.dirty = false;
    c.instances_dirty = true;
  }
  if (!c.scene || c.instances_dirty) {
    c.drop_scene();
    if (gvtb200_scene_create(&c.scene) != GVTB200_OK) engine_error("scene");
    for (const Instance &in : c.instances) {
      auto mesh = mesh_of(in.mesh);
      if (gvtb200_scene_add_instance(c.scene, mesh->device, in.m, in.minv, in.normi, in.lo, in.hi) < 0) engine_error("instance");
    }
    if (gvtb200_scene_commit(c.scene) != GVTB200_OK) engine_error("commit");
    c.instances_dirty = false;
  }
  std::vector<gvtb200_light> ls;
  for (auto &l : c.lights) ls.push_back(l.l);
  if (gvtb200_scene_set_lights(c.scene, ls.data(), (int)ls.size()) != GVTB200_OK) engine_error("lights");
  const CameraRec &cr = cam->second;
  if (gvtb200_scene_set_camera(c.scene, cr.eye, cr.focus, cr.up, cr.fov, cr.depth, cr.samples, cr.jitter) != GVTB200_OK)
    engine_error("camera");
  if (gvtb200_scene_set_film(c.scene, film->second.w, film->second.h) != GVTB200_OK) engine_error("film");
  // adapter::EmbreeStream (Types.h:44): EmbreeStreamMeshAdapter.cpp has no `#define FLAT_SHADING` (EmbreeMeshAdapter.cpp:75
  // does), so its hits are shaded with interpolated vertex normals (:931-946); its explicit minv*ray transform (:308-309) and
  // 1x64 ray streams (CMakeLists.txt:457-458) are what this engine does for every adapter type already.
  const unsigned flags = c.flags | (r.adapter == gvt::render::adapter::EmbreeStream ? (unsigned)GVTB200_FLAG_SMOOTH : 0u);
  if (gvtb200_scene_set_flags(c.scene, flags, c.seed) != GVTB200_OK) engine_error("flags");
  return c.scene;
}


// ---- Tracer<ImageScheduler> / Tracer<DomainScheduler> on several GPUs of this process -------------------------------------
// Renderer.cpp:71-96 picks the tracer from the schedule type; here the tracer is one host thread per GPU (no MPI, no IPC: the
// devices map each other's memory with cudaDeviceEnablePeerAccess) driving the same C ABI calls the one-process-per-GPU
// drivers use:
//   Image / AsyncImage   geometry replicated (ImageTracer.h:111-125 slices the film), every GPU renders its row tiles and
//                        brings them to their place in the host image over its own PCIe link;
//   Domain               instance k lives on GPU k % n (DomainTracer.h:140); exchange rounds with rays stored into the owner's
//                        queue over NVLink, the termination sum (:339-349) formed by the host threads between two barriers;
//   AsyncDomain          the persistent scheduler: one kernel per GPU and frame, no rounds, termination on the device.
// The composited image (root GPU 0) is read back in parallel as well.
struct Barrier { // the threads of one frame meet here (what MPI_Barrier / the gather of DomainTracer.h:339-349 were for)
  std::mutex m;
  std::condition_variable cv;
  int n, waiting = 0, phase = 0;
  explicit Barrier(int n) : n(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const int ph = phase;
    if (++waiting == n) {
      waiting = 0, ++phase;
      cv.notify_all();
    } else
      cv.wait(lk, [&] { return phase != ph; });
  }
};

void multi_build(const RendererRec &r, int n) {
  Context &c = db();
  const CameraRec &cr = c.cameras.at(r.camera);
  const FilmRec &fr = c.films.at(r.film);
  const bool domain = r.schedule == gvt::render::scheduler::Domain || r.schedule == gvt::render::scheduler::AsyncDomain;
  bool mesh_dirty = false;
  for (auto &kv : c.meshes) mesh_dirty |= kv.second->dirty;
  Context::Multi &M = c.multi;
  if (M.n != n || M.schedule != r.schedule || M.w != fr.w || M.h != fr.h || mesh_dirty || c.instances_dirty) {
    c.drop_multi();
    M.n = n, M.schedule = r.schedule, M.w = fr.w, M.h = fr.h;
    M.meshes.assign(n, std::vector<gvtb200_mesh>(c.meshes.size(), nullptr));
    M.scenes.assign(n, nullptr);
    std::vector<std::string> mesh_names;
    for (auto &kv : c.meshes) kv.second->generateNormals(), mesh_names.push_back(kv.first);
    std::vector<int32_t> owner(c.instances.size());
    for (size_t k = 0; k < owner.size(); ++k) owner[k] = domain ? (int32_t)(k % (size_t)n) : 0; // mpiInstanceMap, DomainTracer.h:140
    std::vector<std::string> errs(n);
    std::vector<std::thread> th;
    for (int g = 0; g < n; ++g)
      th.emplace_back([&, g] {
        auto bad = [&](const char *what) { errs[g] = std::string(what) + ": " + gvtb200_last_error(); };
        if (gvtb200_init(g) != GVTB200_OK) return bad("gvtb200_init");
        for (size_t mi = 0; mi < mesh_names.size(); ++mi) { // ImageScheduler: every mesh; DomainScheduler: the meshes of owned instances
          bool need = !domain;
          for (size_t k = 0; k < c.instances.size() && !need; ++k) need = owner[k] == g && c.instances[k].mesh == mesh_names[mi];
          if (!need) continue;
          Mesh &m = *c.meshes[mesh_names[mi]];
          gvtb200_material mm;
          mm.type = m.mat.type, memcpy(mm.kd, m.mat.kd, 12), memcpy(mm.ks, m.mat.ks, 12), mm.alpha = m.mat.alpha;
          std::vector<gvtb200_material> fm(m.faces_to_materials.size());
          const size_t nf = m.faces.size() / 3;
          std::vector<int32_t> fidx;
          for (size_t i = 0; i < fm.size(); ++i) {
            const Material &q = m.faces_to_materials[i];
            fm[i].type = q.type, memcpy(fm[i].kd, q.kd, 12), memcpy(fm[i].ks, q.ks, 12), fm[i].alpha = q.alpha;
          }
          if (!fm.empty()) {
            fidx.resize(nf);
            for (size_t i = 0; i < nf; ++i) fidx[i] = i < fm.size() ? (int32_t)i : -1;
          }
          const bool vc = !m.vertex_colors.empty() && m.vertex_colors.size() == m.vertices.size();
          if (gvtb200_mesh_create(&m.vertices[0].x, (uint32_t)m.vertices.size(), m.faces.data(), (uint32_t)nf,
                                  m.normals.size() == m.vertices.size() ? &m.normals[0].x : nullptr, &mm, fm.empty() ? nullptr : fm.data(),
                                  (uint32_t)fm.size(), fidx.empty() ? nullptr : fidx.data(), vc ? &m.vertex_colors[0].x : nullptr,
                                  &M.meshes[g][mi]) != GVTB200_OK)
            return bad("mesh build");
        }
        if (gvtb200_scene_create(&M.scenes[g]) != GVTB200_OK) return bad("scene");
        for (const Instance &in : c.instances) {
          const size_t mi = std::find(mesh_names.begin(), mesh_names.end(), in.mesh) - mesh_names.begin();
          if (gvtb200_scene_add_instance(M.scenes[g], M.meshes[g][mi], in.m, in.minv, in.normi, in.lo, in.hi) < 0) return bad("instance");
        }
        if (domain && gvtb200_domain_set_owner(M.scenes[g], owner.data(), (int)owner.size(), g, n) != GVTB200_OK) return bad("owner");
      });
    for (auto &t : th) t.join();
    for (auto &e : errs)
      if (!e.empty()) throw std::runtime_error("multi-GPU setup: " + e);
    for (auto &kv : c.meshes) kv.second->dirty = false;
    c.instances_dirty = false;
    if (c.scene) gvtb200_scene_destroy(c.scene), c.scene = nullptr; // the single-GPU scene is rebuilt on demand
    c.instances_dirty = false;
  }
  // per-frame parameters, then (first frame) the commit and the peer mapping
  std::vector<gvtb200_light> ls;
  for (auto &l : c.lights) ls.push_back(l.l);
  const unsigned flags = c.flags | (r.adapter == gvt::render::adapter::EmbreeStream ? (unsigned)GVTB200_FLAG_SMOOTH : 0u);
  for (int g = 0; g < n; ++g) {
    gvtb200_scene sc = M.scenes[g];
    if (gvtb200_scene_set_lights(sc, ls.data(), (int)ls.size()) != GVTB200_OK) engine_error("lights");
    if (gvtb200_scene_set_camera(sc, cr.eye, cr.focus, cr.up, cr.fov, cr.depth, cr.samples, cr.jitter) != GVTB200_OK) engine_error("camera");
    if (gvtb200_scene_set_film(sc, fr.w, fr.h) != GVTB200_OK) engine_error("film");
    if (gvtb200_scene_set_flags(sc, flags, c.seed) != GVTB200_OK) engine_error("flags");
  }
  if (!M.attached) {
    std::vector<gvtb200_peer_desc> descs(n);
    for (int g = 0; g < n; ++g) {
      if (gvtb200_scene_commit(M.scenes[g]) != GVTB200_OK) engine_error("commit");
      if (gvtb200_p2p_export(M.scenes[g], n, &descs[g]) != GVTB200_OK) engine_error("p2p export");
    }
    for (int g = 0; g < n; ++g)
      if (gvtb200_p2p_attach(M.scenes[g], g, n, descs.data()) != GVTB200_OK) engine_error("p2p attach");
    M.attached = true;
  }
}

void render_multi(const RendererRec &r, int n) {
  Context &c = db();
  multi_build(r, n);
  Context::Multi &M = c.multi;
  const FilmRec &fr = c.films.at(r.film);
  c.fb.assign((size_t)fr.w * fr.h * 4, 0.f);
  c.fbw = fr.w, c.fbh = fr.h;
  const int tile_rows = 8;
  Barrier bar(n);
  std::vector<unsigned long long> exported(n, 0);
  std::vector<std::string> errs(n);
  std::vector<std::thread> th;
  for (int g = 0; g < n; ++g)
    th.emplace_back([&, g] {
      gvtb200_scene sc = M.scenes[g];
      bool ok = true;
      auto check = [&](int rc, const char *what) {
        if (rc != GVTB200_OK && errs[g].empty()) errs[g] = std::string(what) + ": " + gvtb200_last_error(), ok = false;
      };
      switch (r.schedule) {
      case gvt::render::scheduler::Domain: { // Tracer<DomainScheduler>::operator(), DomainTracer.h:185-368
        check(gvtb200_domain_begin_frame(sc), "domain_begin_frame");
        check(gvtb200_scene_sync(sc), "sync");
        bar.wait();
        unsigned long long before = 0;
        for (int round = 0; round < 100000; ++round) {
          void *d_total = nullptr;
          check(gvtb200_domain_round_p2p(sc, round, &d_total), "domain_round_p2p");
          check(gvtb200_scene_sync(sc), "sync");
          gvtb200_counters ctr;
          check(gvtb200_scene_counters(sc, &ctr), "counters");
          exported[g] = ok ? ctr.rays_exported : ~0ull;
          bar.wait(); // every rank's round is over: what it stored into its peers is visible
          unsigned long long total = 0;
          bool any_bad = false;
          for (int k = 0; k < n; ++k) total += exported[k], any_bad |= exported[k] == ~0ull;
          bar.wait(); // everybody has read the counts before anybody overwrites them
          if (any_bad || total == before) break; // no ray changed rank in this round: nothing is in flight (DomainTracer.h:339-349)
          before = total;
        }
        if (g != 0) check(gvtb200_domain_composite_p2p(sc, 0), "composite"); // peers first: the root's call waits for them
        bar.wait();
        if (g == 0) check(gvtb200_domain_composite_p2p(sc, 0), "composite");
        check(gvtb200_scene_sync(sc), "sync");
        bar.wait();
        check(gvtb200_framebuffer_read_rows(sc, c.fb.data(), tile_rows, g, n, 0), "read_rows");
        break;
      }
      case gvt::render::scheduler::AsyncDomain: // the persistent scheduler (tracer/Domain/DomainTracer.cpp:109-192 is its relative)
        check(gvtb200_domain_frame_persist(sc, 0, tile_rows), "domain_frame_persist");
        check(gvtb200_scene_sync(sc), "sync");
        bar.wait(); // the root's image is complete (and clamped) when the root's kernel has ended
        check(gvtb200_framebuffer_read_rows(sc, c.fb.data(), tile_rows, g, n, 0), "read_rows");
        break;
      default: // Image, AsyncImage: Tracer<ImageScheduler>::operator(), ImageTracer.h:127-269
        check(gvtb200_render_frame_tiles(sc, 0, fr.h, tile_rows, g, n), "render_frame_tiles");
        check(gvtb200_framebuffer_read_rows(sc, c.fb.data(), tile_rows, g, n, -1), "read_rows");
        break;
      }
      check(gvtb200_scene_sync(sc), "sync");
    });
  for (auto &t : th) t.join();
  for (auto &e : errs)
    if (!e.empty()) throw std::runtime_error("multi-GPU render: " + e);
}

} // namespace

namespace api {

void gvtInit(int, char **, unsigned int) { ensure_init(); } // api.cpp:76-98 (MPI / TBB setup has no counterpart here)

void gvtReset() {
  Context &c = db();
  c.drop_scene();
  c.meshes.clear(), c.instances.clear(), c.lights.clear(), c.cameras.clear(), c.films.clear(), c.renderers.clear(), c.names.clear();
  c.flags = c.seed = 0;
  c.fb.clear(), c.fbw = c.fbh = 0;
}

void createMesh(const std::string name) { // api.cpp:116-123
  db().unique(name);
  db().meshes[name] = std::make_shared<Mesh>();
}
void addMeshVertices(const std::string name, const unsigned &n, const float *vertices, const bool tessellate, const std::string) {
  auto m = mesh_of(name);
  if (tessellate) throw std::runtime_error("addMeshVertices: tessellate=true (qhull) is not supported by this engine");
  for (unsigned i = 0; i < n * 3; i += 3) m->vertices.push_back(V3{ vertices[i], vertices[i + 1], vertices[i + 2] });
  m->dirty = true;
}
void addMeshTriangles(const std::string name, const unsigned &n, const unsigned *triangles) { // api.cpp:173-180
  auto m = mesh_of(name);
  for (unsigned i = 0; i < n * 3; i += 3) m->addFace((int)triangles[i], (int)triangles[i + 1], (int)triangles[i + 2]);
  m->dirty = true;
}
void addMeshTriangles0(const std::string name, const unsigned &n, const unsigned *t) { // ObjReader.cpp:192-194
  auto m = mesh_of(name);
  for (unsigned i = 0; i < n * 3; ++i) {
    if (t[i] >= m->vertices.size()) throw std::runtime_error("Vertex index outside bounds : " + std::to_string(t[i]));
    m->faces.push_back(t[i]);
  }
  m->dirty = true;
}
void addMeshFaceNormals(const std::string name, const unsigned &n, const float *normals) {
  auto m = mesh_of(name);
  for (unsigned i = 0; i < n * 3; i += 3) m->face_normals.push_back(V3{ normals[i], normals[i + 1], normals[i + 2] });
}
void addMeshVertexNormals(const std::string name, const unsigned &n, const float *normals) {
  auto m = mesh_of(name);
  for (unsigned i = 0; i < n * 3; i += 3) m->normals.push_back(V3{ normals[i], normals[i + 1], normals[i + 2] });
  m->dirty = true;
}
void finishMesh(const std::string name, const bool compute_normal) { // api.cpp:182-192
  auto m = mesh_of(name);
  m->computeBoundingBox();
  if (compute_normal) m->generateNormals();
}
void addMeshMaterial(const std::string name, const unsigned mattype, const float *kd, const float alpha) {
  auto m = mesh_of(name);
  m->mat = Material();
  m->mat.type = (int)mattype, memcpy(m->mat.kd, kd, 12), m->mat.alpha = alpha;
  m->dirty = true;
}
void addMeshMaterial(const std::string name, const unsigned mattype, const float *kd, const float *ks, const float alpha) {
  auto m = mesh_of(name);
  m->mat = Material();
  m->mat.type = (int)mattype, memcpy(m->mat.kd, kd, 12), memcpy(m->mat.ks, ks, 12), m->mat.alpha = alpha;
  m->dirty = true;
}
void addMeshMaterials(const std::string name, const unsigned n, const unsigned *mattype, const float *kd, const float *ks,
                      const float *alpha) { // api.cpp:250-264: one material per face
  auto m = mesh_of(name);
  for (unsigned i = 0; i < n; ++i) {
    Material mat;
    mat.type = (int)mattype[i], memcpy(mat.kd, kd + 3 * i, 12), memcpy(mat.ks, ks + 3 * i, 12), mat.alpha = alpha[i];
    m->faces_to_materials.push_back(mat);
  }
  m->dirty = true;
}
void addMeshVertexColor(const std::string name, const unsigned n, const float *kd) {
  auto m = mesh_of(name);
  for (unsigned i = 0; i < n; ++i) m->vertex_colors.push_back(V3{ kd[3 * i], kd[3 * i + 1], kd[3 * i + 2] });
  m->dirty = true;
}

void addInstance(std::string instname, std::string meshname, const float *am) { // api.cpp:292-322
  auto mesh = mesh_of(meshname);
  db().unique(instname);
  Instance in;
  in.name = instname, in.mesh = meshname;
  memcpy(in.m, am, 64); // glm::make_mat4: column-major
  mat4_inverse(in.m, in.minv);
  mat3_inverse_transpose(in.m, in.normi);
  float il[3], ih[3];
  mat4_point(in.m, mesh->bmin, il);
  mat4_point(in.m, mesh->bmax, ih);
  for (int i = 0; i < 3; ++i) in.lo[i] = fastmin(il[i], ih[i]), in.hi[i] = fastmax(il[i], ih[i]); // Box3D(vmin,vmax) BBox.cpp:40-46
  db().instances.push_back(in);
  db().instances_dirty = true;
}

void addPointLight(std::string name, const float *pos, const float *color) {
  db().unique(name);
  LightRec l;
  l.name = name;
  memset(&l.l, 0, sizeof(l.l));
  l.l.type = 0, memcpy(l.l.position, pos, 12), memcpy(l.l.color, color, 12);
  db().lights.push_back(l);
}
void addAreaLight(std::string name, const float *pos, const float *color, const float *n, float w, float h) {
  db().unique(name);
  LightRec l;
  l.name = name;
  memset(&l.l, 0, sizeof(l.l));
  l.l.type = 2, memcpy(l.l.position, pos, 12), memcpy(l.l.color, color, 12), memcpy(l.l.normal, n, 12), l.l.width = w, l.l.height = h;
  db().lights.push_back(l);
}
void modifyLight(std::string name, const float *pos, const float *color) {
  for (auto &l : db().lights)
    if (l.name == name) memcpy(l.l.position, pos, 12), memcpy(l.l.color, color, 12);
}
void modifyLight(std::string name, const float *pos, const float *color, const float *n, float w, float h) {
  for (auto &l : db().lights)
    if (l.name == name)
      memcpy(l.l.position, pos, 12), memcpy(l.l.color, color, 12), memcpy(l.l.normal, n, 12), l.l.width = w, l.l.height = h;
}
void addCamera(std::string name, const float *pos, const float *focus, const float *up, float fov, int depth, int samples, float jitter) {
  db().unique(name);
  CameraRec c;
  memcpy(c.eye, pos, 12), memcpy(c.focus, focus, 12), memcpy(c.up, up, 12);
  c.fov = fov, c.depth = depth, c.samples = samples, c.jitter = jitter;
  db().cameras[name] = c;
}
void modifyCamera(std::string name, const float *pos, const float *focus, const float *up, float fov, int depth, int samples,
                  float jitter) {
  auto it = db().cameras.find(name);
  if (it == db().cameras.end()) return;
  CameraRec &c = it->second;
  memcpy(c.eye, pos, 12), memcpy(c.focus, focus, 12), memcpy(c.up, up, 12);
  c.fov = fov, c.depth = depth, c.samples = samples, c.jitter = jitter;
}
void modifyCamera(std::string name, const float *pos, const float *focus, const float *up, float fov) {
  auto it = db().cameras.find(name);
  if (it == db().cameras.end()) return;
  CameraRec &c = it->second;
  memcpy(c.eye, pos, 12), memcpy(c.focus, focus, 12), memcpy(c.up, up, 12);
  c.fov = fov;
}
void addFilm(std::string name, int w, int h, std::string path) {
  db().unique(name);
  db().films[name] = FilmRec{ w, h, path };
}
void modifyFilm(std::string name, int w, int h, std::string path) {
  auto it = db().films.find(name);
  if (it == db().films.end()) return;
  it->second = FilmRec{ w, h, path };
}
void addRenderer(std::string name, int adapter, int schedule, std::string const &Camera, std::string const &Film, bool volume) {
  db().unique(name);
  db().renderers[name] = RendererRec{ adapter, schedule, Camera, Film, volume };
}
void modifyRenderer(std::string name, int adapter, int schedule, std::string const &Camera, std::string const &Film) {
  auto it = db().renderers.find(name);
  if (it == db().renderers.end()) return;
  it->second.adapter = adapter, it->second.schedule = schedule, it->second.camera = Camera, it->second.film = Film;
}

void *engineScene(std::string name) {
  auto it = db().renderers.find(name);
  if (it == db().renderers.end()) throw std::runtime_error("Suplied renderer " + name + " is not valid"); // Renderer.cpp:43
  return (void *)device_scene(it->second);
}

// gvtRenderer::render (Renderer.cpp:106-115): camera rays + Tracer<...>::operator() + composite, all on the device.
// On one GPU a single rank owns every instance, so scheduler::Domain degenerates to the image path exactly as one MPI rank
// does in the reference; with api::setGpus(n > 1) the schedule type picks the multi-GPU tracer (render_multi above).
void render(std::string name) {
  auto it = db().renderers.find(name);
  if (it == db().renderers.end()) throw std::runtime_error("Suplied renderer " + name + " is not valid");
  if (it->second.volume) throw std::runtime_error("volume rendering is not supported by this engine");
  if (db().gpus > 1) return render_multi(it->second, db().gpus);
  gvtb200_scene s = device_scene(it->second);
  Context &c = db();
  const FilmRec &f = c.films[it->second.film];
  c.fb.resize((size_t)f.w * f.h * 4);
  c.fbw = f.w, c.fbh = f.h;
  if (gvtb200_render_frame(s, 0, f.h, c.fb.data(), nullptr) != GVTB200_OK) engine_error("render");
}
void gvtsync() {} // cntx::rcontext::sync broadcasts dirty nodes between MPI ranks (context.h:381-452): single process here

const std::vector<float> &framebuffer(int *width, int *height) {
  if (width) *width = db().fbw;
  if (height) *height = db().fbh;
  return db().fb;
}
void setEngineFlags(unsigned flags, unsigned seed) { db().flags = flags, db().seed = seed; }
void setGpus(int n) {
  if (n < 1 || n > 16) throw std::runtime_error("setGpus: 1..16 GPUs");
  if (n != db().gpus) db().drop_multi();
  db().gpus = n;
}

// IceTComposite::write (composite/IceTComposite.cpp:119-158): P6, rows bottom-to-top, (unsigned char)(c*255)
void writeimage(std::string name, std::string output) {
  Context &c = db();
  if (output.empty()) {
    auto r = c.renderers.find(name);
    if (r != c.renderers.end() && c.films.count(r->second.film)) output = c.films[r->second.film].path;
  }
  std::stringstream header;
  header << "P6" << std::endl << c.fbw << " " << c.fbh << std::endl << "255" << std::endl;
  std::fstream file;
  file.open((output + ".ppm").c_str(), std::fstream::out | std::fstream::trunc | std::fstream::binary);
  if (!file) throw std::runtime_error("cannot open " + output + ".ppm");
  file << header.str();
  std::vector<unsigned char> row(3 * (size_t)c.fbw);
  for (int j = c.fbh - 1; j >= 0; j--) {
    for (int i = 0; i < c.fbw; ++i) {
      size_t index = 4 * ((size_t)j * c.fbw + i);
      row[3 * i + 0] = (unsigned char)(c.fb[index + 0] * 255);
      row[3 * i + 1] = (unsigned char)(c.fb[index + 1] * 255);
      row[3 * i + 2] = (unsigned char)(c.fb[index + 2] * 255);
    }
    file.write((const char *)row.data(), row.size());
  }
  file.close();
}

} // namespace api

// ---- C entry points over the C++ API (ctypes / cgo / JNI friendly); 0 = ok, -1 = exception (message kept) ----------
namespace {
thread_local std::string g_api_err;
template <class F> int guard(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_api_err = e.what();
    return -1;
  }
}
} // namespace
#define GAPI extern "C" __attribute__((visibility("default")))
GAPI const char *gvtapi_last_error() { return g_api_err.c_str(); }
GAPI int gvtapi_init() { return guard([&] { api::gvtInit(0, nullptr); }); }
GAPI int gvtapi_reset() { return guard([&] { api::gvtReset(); }); }
GAPI int gvtapi_create_mesh(const char *n) { return guard([&] { api::createMesh(n); }); }
GAPI int gvtapi_add_mesh_vertices(const char *n, unsigned k, const float *v) { return guard([&] { api::addMeshVertices(n, k, v); }); }
GAPI int gvtapi_add_mesh_triangles(const char *n, unsigned k, const unsigned *t) { return guard([&] { api::addMeshTriangles(n, k, t); }); }
GAPI int gvtapi_add_mesh_triangles0(const char *n, unsigned k, const unsigned *t) { return guard([&] { api::addMeshTriangles0(n, k, t); }); }
GAPI int gvtapi_add_mesh_vertex_normals(const char *n, unsigned k, const float *v) { return guard([&] { api::addMeshVertexNormals(n, k, v); }); }
GAPI int gvtapi_finish_mesh(const char *n, int compute_normal) { return guard([&] { api::finishMesh(n, compute_normal != 0); }); }
GAPI int gvtapi_add_mesh_material(const char *n, unsigned type, const float *kd, const float *ks, float alpha) {
  return guard([&] {
    if (ks)
      api::addMeshMaterial(n, type, kd, ks, alpha);
    else
      api::addMeshMaterial(n, type, kd, alpha);
  });
}
GAPI int gvtapi_add_mesh_materials(const char *n, unsigned k, const unsigned *type, const float *kd, const float *ks, const float *alpha) {
  return guard([&] { api::addMeshMaterials(n, k, type, kd, ks, alpha); });
}
GAPI int gvtapi_add_mesh_vertex_color(const char *n, unsigned k, const float *c) { return guard([&] { api::addMeshVertexColor(n, k, c); }); }
GAPI int gvtapi_add_instance(const char *inst, const char *mesh, const float *m) { return guard([&] { api::addInstance(inst, mesh, m); }); }
GAPI int gvtapi_add_point_light(const char *n, const float *p, const float *c) { return guard([&] { api::addPointLight(n, p, c); }); }
GAPI int gvtapi_add_area_light(const char *n, const float *p, const float *c, const float *nn, float w, float h) {
  return guard([&] { api::addAreaLight(n, p, c, nn, w, h); });
}
GAPI int gvtapi_modify_light(const char *n, const float *p, const float *c) { return guard([&] { api::modifyLight(n, p, c); }); }
GAPI int gvtapi_modify_area_light(const char *n, const float *p, const float *c, const float *nn, float w, float h) {
  return guard([&] { api::modifyLight(n, p, c, nn, w, h); });
}
GAPI int gvtapi_add_mesh_face_normals(const char *n, unsigned k, const float *v) { return guard([&] { api::addMeshFaceNormals(n, k, v); }); }
GAPI int gvtapi_modify_renderer(const char *n, int adapter, int schedule, const char *cam, const char *film) {
  return guard([&] { api::modifyRenderer(n, adapter, schedule, cam, film); });
}
GAPI int gvtapi_add_renderer_volume(const char *n, int adapter, int schedule, const char *cam, const char *film, int volume) {
  return guard([&] { api::addRenderer(n, adapter, schedule, cam, film, volume != 0); });
}
GAPI int gvtapi_modify_camera_pose(const char *n, const float *p, const float *f, const float *u, float fov) {
  return guard([&] { api::modifyCamera(n, p, f, u, fov); });
}
GAPI int gvtapi_gvtsync() { return guard([&] { api::gvtsync(); }); }
GAPI int gvtapi_add_camera(const char *n, const float *p, const float *f, const float *u, float fov, int depth, int samples, float jitter) {
  return guard([&] { api::addCamera(n, p, f, u, fov, depth, samples, jitter); });
}
GAPI int gvtapi_modify_camera(const char *n, const float *p, const float *f, const float *u, float fov, int depth, int samples, float jitter) {
  return guard([&] { api::modifyCamera(n, p, f, u, fov, depth, samples, jitter); });
}
GAPI int gvtapi_add_film(const char *n, int w, int h, const char *path) { return guard([&] { api::addFilm(n, w, h, path); }); }
GAPI int gvtapi_modify_film(const char *n, int w, int h, const char *path) { return guard([&] { api::modifyFilm(n, w, h, path); }); }
GAPI int gvtapi_add_renderer(const char *n, int adapter, int schedule, const char *cam, const char *film) {
  return guard([&] { api::addRenderer(n, adapter, schedule, cam, film); });
}
GAPI int gvtapi_render(const char *n) { return guard([&] { api::render(n); }); }
GAPI int gvtapi_set_gpus(int n) { return guard([&] { api::setGpus(n); }); }
GAPI int gvtapi_writeimage(const char *n, const char *out) { return guard([&] { api::writeimage(n, out); }); }
GAPI int gvtapi_set_engine_flags(unsigned flags, unsigned seed) { return guard([&] { api::setEngineFlags(flags, seed); }); }
GAPI int gvtapi_framebuffer(float *rgba, size_t n_floats, int *w, int *h) {
  return guard([&] {
    const std::vector<float> &fb = api::framebuffer(w, h);
    if (rgba) {
      if (n_floats < fb.size()) throw std::runtime_error("framebuffer: buffer too small");
      memcpy(rgba, fb.data(), fb.size() * sizeof(float));
    }
  });
}
GAPI void *gvtapi_engine_scene(const char *renderer) {
  void *p = nullptr;
  guard([&] { p = api::engineScene(renderer); });
  return p;
}
// what api::addInstance stored for instance i (parity hook): m, minv (16), normi (9), box lo/hi (3+3)
GAPI int gvtapi_instance_get(int i, float *m, float *minv, float *normi, float *box) {
  return guard([&] {
    if (i < 0 || i >= (int)db().instances.size()) throw std::runtime_error("no such instance");
    const Instance &in = db().instances[i];
    if (m) memcpy(m, in.m, 64);
    if (minv) memcpy(minv, in.minv, 64);
    if (normi) memcpy(normi, in.normi, 36);
    if (box) memcpy(box, in.lo, 12), memcpy(box + 3, in.hi, 12);
  });
}
// Mesh state after finishMesh (parity hook): number of faces kept, faces (0-based), vertex normals, bbox
GAPI int gvtapi_mesh_get(const char *n, unsigned *n_faces, unsigned *faces, float *normals, float *bbox) {
  return guard([&] {
    auto m = mesh_of(n);
    if (n_faces) *n_faces = (unsigned)(m->faces.size() / 3);
    if (faces) memcpy(faces, m->faces.data(), m->faces.size() * 4);
    if (normals && !m->normals.empty()) memcpy(normals, m->normals.data(), m->normals.size() * 12);
    if (bbox) memcpy(bbox, m->bmin, 12), memcpy(bbox + 3, m->bmax, 12);
  });
}
