// B200MeshAdapter.h -- the reference-side binding: a gvt::render::Adapter that forwards to libgvtb200.so.
//
// This file is meant to be dropped into a GraviT checkout as src/gvt/render/adapter/b200/B200MeshAdapter.h; it
// includes GraviT's own headers (it does not build inside this repo on its own) and replaces
// adapter/embree/EmbreeMeshAdapter.{h,cpp} for the path Adapter::trace (Adapter.h:82-84).  INTEGRATION.md lists the
// three lines of tracer glue.  tests/test_host_cpu.py syntax-checks it against /root/reference when that is present.
#ifndef GVT_RENDER_ADAPTER_B200_MESH_ADAPTER_H
#define GVT_RENDER_ADAPTER_B200_MESH_ADAPTER_H

#include <gvt/render/Adapter.h>
#include <gvt/render/data/Primitives.h>
#include <gvt/render/data/scene/Light.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <tuple>
#include <vector>

#include "gvtb200.h"

namespace gvt {
namespace render {
namespace adapter {
namespace b200 {
namespace data {

class B200MeshAdapter : public gvt::render::Adapter {
public:
  // EmbreeMeshAdapter::EmbreeMeshAdapter (EmbreeMeshAdapter.cpp:125-162): normals + BVH build, once per mesh
  B200MeshAdapter(std::shared_ptr<gvt::render::data::primitives::Data> d) : Adapter(d), handle(nullptr) {
    static_assert(sizeof(gvt::render::actor::Ray) == sizeof(gvtb200_ray), "Ray record layout");
    auto mesh = std::dynamic_pointer_cast<gvt::render::data::primitives::Mesh>(d);
    if (!mesh) throw std::runtime_error("B200MeshAdapter: data node is not a Mesh");
    mesh->generateNormals();
    static bool inited = false;
    if (!inited) {
      if (gvtb200_init(0) != GVTB200_OK) throw std::runtime_error(gvtb200_last_error());
      inited = true;
    }
    std::vector<uint32_t> faces(3 * mesh->faces.size());
    for (size_t i = 0; i < mesh->faces.size(); ++i) {
      faces[3 * i + 0] = (uint32_t)std::get<0>(mesh->faces[i]);
      faces[3 * i + 1] = (uint32_t)std::get<1>(mesh->faces[i]);
      faces[3 * i + 2] = (uint32_t)std::get<2>(mesh->faces[i]);
    }
    auto pod = [](const gvt::render::data::primitives::Material *m) {
      gvtb200_material p;
      p.type = m->type;
      for (int k = 0; k < 3; ++k) p.kd[k] = m->kd[k], p.ks[k] = m->ks[k];
      p.alpha = m->alpha;
      return p;
    };
    gvtb200_material mm = pod(mesh->mat);
    std::vector<gvtb200_material> fm;
    std::vector<int32_t> fidx;
    if (!mesh->faces_to_materials.empty()) { // EmbreeMeshAdapter.cpp:565-566
      fidx.assign(mesh->faces.size(), -1);
      for (size_t i = 0; i < mesh->faces.size() && i < mesh->faces_to_materials.size(); ++i)
        if (mesh->faces_to_materials[i]) {
          fidx[i] = (int32_t)fm.size();
          fm.push_back(pod(mesh->faces_to_materials[i]));
        }
    }
    const bool vc = mesh->vertex_colors.size() == mesh->vertices.size() && !mesh->vertex_colors.empty();
    // glm::vec3 is three packed floats: the vectors are passed as they lie
    if (gvtb200_mesh_create(&mesh->vertices[0][0], (uint32_t)mesh->vertices.size(), faces.data(), (uint32_t)mesh->faces.size(),
                            mesh->normals.size() == mesh->vertices.size() ? &mesh->normals[0][0] : nullptr, &mm,
                            fm.empty() ? nullptr : fm.data(), (uint32_t)fm.size(), fidx.empty() ? nullptr : fidx.data(),
                            vc ? &mesh->vertex_colors[0][0] : nullptr, &handle) != GVTB200_OK)
      throw std::runtime_error(gvtb200_last_error());
  }
  virtual ~B200MeshAdapter() { gvtb200_mesh_destroy(handle); }

  // EmbreeMeshAdapter::trace (EmbreeMeshAdapter.cpp:625-660)
  virtual void trace(gvt::render::actor::RayVector &rayList, gvt::render::actor::RayVector &moved_rays, glm::mat4 *m, glm::mat4 *minv,
                     glm::mat3 *normi, std::vector<std::shared_ptr<gvt::render::data::scene::Light> > &lights, size_t begin = 0,
                     size_t end = 0) {
    using namespace gvt::render::data::scene;
    if (end == 0) end = rayList.size();
    std::vector<gvtb200_light> ls;
    for (auto &l : lights) {
      gvtb200_light p;
      memset(&p, 0, sizeof(p));
      for (int k = 0; k < 3; ++k) p.position[k] = l->position[k];
      if (l->LightT == Light::Area) {
        auto *a = static_cast<AreaLight *>(l.get());
        p.type = 2;
        for (int k = 0; k < 3; ++k) p.color[k] = a->color[k], p.normal[k] = a->LightNormal[k];
        p.width = a->LightWidth, p.height = a->LightHeight;
      } else if (auto *pl = dynamic_cast<PointLight *>(l.get())) {
        p.type = 0;
        for (int k = 0; k < 3; ++k) p.color[k] = pl->color[k];
      } else if (auto *al = dynamic_cast<AmbientLight *>(l.get())) {
        p.type = 1;
        for (int k = 0; k < 3; ++k) p.color[k] = al->color[k];
      }
      ls.push_back(p);
    }
    // every input ray can come back, plus one shadow ray per light per bounce level (EmbreeMeshAdapter.cpp:584-601 loops while
    // depth > 0): n * (1 + lights * max depth).  Should the estimate still be short, gvtb200_trace reports GVTB200_ERR_CAPACITY
    // with rayList untouched and the call is repeated with a larger buffer.
    const size_t n = end - begin;
    int max_depth = 1;
    for (size_t i = begin; i < end; ++i) max_depth = std::max(max_depth, std::min(rayList[i].mice.depth, 64));
    size_t cap = n * (1 + ls.size() * (size_t)max_depth) + 16;
    std::vector<gvt::render::actor::Ray> out;
    size_t n_moved = 0;
    for (int attempt = 0;; ++attempt) {
      out.resize(cap);
      n_moved = 0;
      int rc = gvtb200_trace(handle, reinterpret_cast<gvtb200_ray *>(rayList.data()), rayList.size(), begin, end,
                             reinterpret_cast<gvtb200_ray *>(out.data()), out.size(), &n_moved, &(*m)[0][0], &(*minv)[0][0],
                             &(*normi)[0][0], ls.data(), (int)ls.size(), /*flags: HEAD = flat shading*/ 0u, /*seed*/ 0u);
      if (rc == GVTB200_OK) break;
      if (rc != GVTB200_ERR_CAPACITY || attempt >= 4) throw std::runtime_error(gvtb200_last_error());
      cap *= 2;
    }
    std::lock_guard<std::mutex> lock(_outqueue); // EmbreeMeshAdapter.cpp:619-621
    moved_rays.insert(moved_rays.end(), out.begin(), out.begin() + n_moved);
  }

protected:
  gvtb200_mesh handle;
};

} // namespace data
} // namespace b200
} // namespace adapter
} // namespace render
} // namespace gvt
#endif
