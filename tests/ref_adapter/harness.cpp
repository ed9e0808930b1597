// harness.cpp -- TEST INFRASTRUCTURE: runs the reference-side binding (gravit_b200/host/gvt_shim/B200MeshAdapter.h) for real.
//
// This file is ours.  It #includes GraviT's own headers from where they lie (/root/reference/src, third-party/glm) and is
// linked against oracle/_ref/libgvt_ref.so (the reference's Ray.cpp, BBox.cpp, Light.cpp, Mesh.cpp, Material.cpp compiled in
// place by oracle/Makefile) and against libgvtb200.so.  It builds a REAL gvt::render::data::primitives::Mesh, a REAL
// gvt::render::actor::RayVector and a REAL PointLight, constructs the adapter exactly as a GraviT tracer would
// (algorithm/ImageTracer.h:184-232: one adapter per mesh) and calls Adapter::trace (Adapter.h:82-84).  The caller
// (tests/test_gpu_parity.py) compares rayList and moved_rays with the oracle byte for byte.  Built into oracle/_ref/ (no
// reference source is copied into this repo; the output never ships with the product).
#include "B200MeshAdapter.h"

#include <cstring>
#include <memory>

using namespace gvt::render::actor;
using namespace gvt::render::data::primitives;
using namespace gvt::render::data::scene;

#define H_API extern "C" __attribute__((visibility("default")))

static std::string g_err;
H_API const char *refadapter_last_error() { return g_err.c_str(); }

// returns the number of moved rays (appended after n_moved_in pre-existing ones, as the tracer's vector would hold them),
// or -1 on an exception.  rays80 is updated in place (Adapter::trace mutates rayList).
H_API long refadapter_trace(const float *verts, int nv, const unsigned *tris1, int nt, const float *kd, unsigned char *rays80, long n_rays,
                            long begin, long end, const float *m16, const float *light_pos, const float *light_color, int n_lights,
                            unsigned char *moved80, long moved_cap) {
  try {
    Material *mat = new Material(); // Material.h:59-75 defaults (Lambert)
    mat->kd = glm::vec3(kd[0], kd[1], kd[2]);
    auto mesh = std::make_shared<Mesh>(mat);
    for (int i = 0; i < nv; ++i) mesh->addVertex(glm::vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
    for (int i = 0; i < nt; ++i) mesh->addFace((int)tris1[3 * i], (int)tris1[3 * i + 1], (int)tris1[3 * i + 2]); // 1-based, filtered
    mesh->computeBoundingBox();
    gvt::render::adapter::b200::data::B200MeshAdapter adapter(mesh); // generateNormals + BLAS build on the GPU
    RayVector rayList((size_t)n_rays), moved;
    for (long i = 0; i < n_rays; ++i) rayList[(size_t)i] = Ray(rays80 + 80 * i);
    moved.reserve((size_t)n_rays * 10); // ImageTracer.h:240
    glm::mat4 m = glm::make_mat4(m16), minv = glm::inverse(m);
    glm::mat3 normi = glm::transpose(glm::inverse(glm::mat3(m))); // api.cpp:300-308
    std::vector<std::shared_ptr<Light>> lights;
    for (int l = 0; l < n_lights; ++l)
      lights.push_back(std::make_shared<PointLight>(glm::vec3(light_pos[3 * l], light_pos[3 * l + 1], light_pos[3 * l + 2]),
                                                    glm::vec3(light_color[3 * l], light_color[3 * l + 1], light_color[3 * l + 2])));
    gvt::render::Adapter *base = &adapter; // through the base class, as the tracers call it
    base->trace(rayList, moved, &m, &minv, &normi, lights, (size_t)begin, (size_t)end);
    for (long i = 0; i < n_rays; ++i) memcpy(rays80 + 80 * i, &rayList[(size_t)i], 80);
    if ((long)moved.size() > moved_cap) {
      g_err = "moved buffer too small";
      return -1;
    }
    for (size_t i = 0; i < moved.size(); ++i) memcpy(moved80 + 80 * i, &moved[i], 80);
    return (long)moved.size();
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}
