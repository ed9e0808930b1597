// readers.h -- scene ingestion for the host mirror: the on-disk formats either side of the render path.
//
//   gvt::render::data::domain::reader::ObjReader   src/gvt/render/data/reader/ObjReader.{h,cpp}: Wavefront OBJ + MTL
//       through tinyobjloader (third-party/tinyobjloader-master/tiny_obj_loader.cc), flattened shape by shape
//   gvt::render::data::domain::reader::PlyReader   src/gvt/render/data/reader/PlyReader.{h,cpp} (Greg Turk's ply.c,
//       third-party/ply/ply.c): a directory of .ply blocks -> meshes "Mesh0", "Mesh1", ... through the api:: calls
//   gvtapps::render::ConfigFileLoader              src/apps/render/ConfigFileLoader.{h,cpp}: data/*.conf scene files
//
// Same class names, constructor arguments and results as the reference.  The readers produce host arrays and hand them
// to namespace api (api.h); nothing here touches the device.  File errors throw std::runtime_error where the reference
// calls exit() (ObjReader.cpp:135-138, PlyReader.cpp:80-93).
#ifndef GVTB200_HOST_READERS_H
#define GVTB200_HOST_READERS_H
#include <string>
#include <vector>

namespace gvt {
namespace render {
namespace data {
namespace primitives {
enum MaterialType { LAMBERT = 0, PHONG = 1, BLINN = 2 }; // data/primitives/Material.h:41-57 (the three Shade() handles)
}
namespace domain {
namespace reader {

// what the reference keeps in a primitives::Mesh after a reader ran (data/primitives/Mesh.h:63-103)
struct MeshMaterial { // primitives::Material, Material.h:59-89
  int type = 0;
  float ka[3] = { 0.f, 0.f, 0.f }, kd[3] = { .5f, .5f, .5f }, ks[3] = { .5f, .5f, .5f };
  float alpha = 1.f;
};
struct MeshData {
  std::vector<float> vertices;         // xyz
  std::vector<float> normals;          // xyz per vertex as found in the file (normalised); apps regenerate them
  std::vector<float> vertex_colors;    // rgb per vertex in [0,1] (PLY red/green/blue), empty when absent
  std::vector<unsigned> faces;         // 0-based, three per face
  std::vector<MeshMaterial> materials; // the MTL library (empty: one default material)
  std::vector<int> faces_to_materials; // index into materials per face (empty when materials is)
  float bounds_min[3], bounds_max[3];
};

class ObjReader {
public:
  // ObjReader.cpp:63: parses the file at once; every material of the MTL library gets `material_type`
  ObjReader(const std::string filename = "", int material_type = gvt::render::data::primitives::LAMBERT);
  virtual ~ObjReader();
  MeshData *getMesh() { return &objMesh; }
  // what SimpleFileLoadApp.cpp:137-168 does with the mesh, expressed through the public scene API:
  // createMesh, addMeshVertices, faces as stored (0-based, unfiltered), the materials, finishMesh(compute_normal = true)
  void addToScene(const std::string &meshname) const;

private:
  MeshData objMesh;
};

class PlyReader {
public:
  // PlyReader.cpp:54: `rootdir` is a directory; every *.ply in it (glob order) becomes a mesh "Mesh<k>" via api::createMesh,
  // addMeshVertices, addMeshTriangles (1-based, degenerate faces dropped), addMeshVertexColor when the vertex element
  // has more than five properties, finishMesh.  `dist` is accepted and unused, as in the reference.
  PlyReader(const std::string rootdir, bool dist = false);
  virtual ~PlyReader();
  const std::vector<std::string> &getMeshes() const { return meshes; }
  // one identity-transform instance "inst<k>" per mesh (PlyApp.cpp:160-184)
  void addInstances() const;
  // one file, no scene calls (ascii, binary_little_endian, binary_big_endian)
  static void readFile(const std::string &path, MeshData &out);

private:
  std::vector<std::string> meshes;
};

} // namespace reader
} // namespace domain
} // namespace data
} // namespace render
} // namespace gvt

namespace gvtapps {
namespace render {

// data/*.conf in the keyed format ConfigFileLoader.cpp:95-283 parses:
//   F w h | C eye(3) focus(3) up(3) fov_degrees | G file.obj translate(3) rotate(3) scale(3) | LP pos(3) color(3) |
//   LA pos(3) normal(3) width height color(3) | DT MANTA|OPTIX|EMBREE|B200 | ST IMAGE|DOMAIN|HYBRID | AT NONE|BVH | RT n
// Lines that start with '#' and unknown keys are skipped.  The loader issues the api:: calls that build the scene:
// mesh "Mesh<k>" per distinct OBJ path, instance "inst<k>" per G line, lights "PointLight<k>" / "AreaLight<k>", camera
// "Camera", film "Film", renderer "Renderer" with the schedule of the ST line.
class ConfigFileLoader {
public:
  enum AccelType { NoAccel, BVH };
  ConfigFileLoader(const std::string filename = "");
  virtual ~ConfigFileLoader();
  int scheduler_type = 0;   // 0 image, 1 domain, 2 hybrid (ConfigFileLoader.h:85)
  int domain_type = 0;      // 0 manta, 1 optix, 2 embree (ConfigFileLoader.h:87); 3 = B200
  AccelType accel_type = NoAccel;
  int width = 0, height = 0;
  int n_instances = 0, n_lights = 0;
};

} // namespace render
} // namespace gvtapps
#endif
