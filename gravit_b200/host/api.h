// api.h -- host-side mirror of GraviT's public scene API (reference: src/gvt/render/api/api.h:26-263) on top of the
// B200 C ABI (include/gvtb200.h).  Same namespace, names, argument meaning and error behaviour as the reference:
// arrays are borrowed and copied during the call (api.cpp:138-139,177-179); unknown names throw std::runtime_error
// (core/cntx/context.h:89,98,241); modify* calls on unknown names silently return (api.cpp:364-366); faces are
// 1-based and degenerate faces are dropped (data/primitives/Mesh.cpp:102-112).  The context database is a
// single-process in-memory stand-in (cntx is out of scope as a system).
//
// Not mirrored (out of scope): tessellate=true (qhull), volumes.
#ifndef GVTB200_HOST_API_H
#define GVTB200_HOST_API_H
#include <string>
#include <thread>
#include <vector>

namespace gvt {
namespace render {
namespace adapter {
// Types.h:33-46 plus the new engine.  Every enumerator selects the B200 engine: there is no CPU adapter here.
enum RenderType { Surface, Manta, Optix, Embree, Ospray, Pvol, EmbreeStream, Heterogeneous, B200 };
} // namespace adapter
namespace scheduler {
enum ScheduleType { Image, Domain, AsyncImage, AsyncDomain }; // Types.h:48-62 (hybrid schedules are compiled out at HEAD)
} // namespace scheduler
} // namespace render
} // namespace gvt

namespace api {

void gvtInit(int argc, char **argv, unsigned int threads = std::thread::hardware_concurrency());

void createMesh(const std::string name);
void addMeshVertices(const std::string name, const unsigned &n, const float *vertices, const bool tessellate = false,
                     const std::string qhullargs = std::string());
void addMeshTriangles(const std::string name, const unsigned &n, const unsigned *triangles);
void addMeshFaceNormals(const std::string name, const unsigned &n, const float *normals);
void addMeshVertexNormals(const std::string name, const unsigned &n, const float *normals);
void finishMesh(const std::string name, const bool compute_normal = true);
void addMeshMaterial(const std::string name, const unsigned mattype, const float *kd, const float alpha = 1.f);
void addMeshMaterial(const std::string name, const unsigned mattype, const float *kd, const float *ks, const float alpha = 1.f);
void addMeshMaterials(const std::string name, const unsigned n, const unsigned *mattype, const float *kd, const float *ks,
                      const float *alpha);
void addMeshVertexColor(const std::string name, const unsigned n, const float *kd);
void addInstance(std::string instname, std::string meshname, const float *m);
void addPointLight(std::string name, const float *pos, const float *color);
void addAreaLight(std::string name, const float *pos, const float *color, const float *n, float w, float h);
void modifyLight(std::string name, const float *pos, const float *color);
void modifyLight(std::string name, const float *pos, const float *color, const float *n, float w, float h);
void addCamera(std::string name, const float *pos, const float *focus, const float *up, float fov, int depth, int samples,
               float jitter);
void modifyCamera(std::string name, const float *pos, const float *focus, const float *up, float fov, int depth, int samples,
                  float jitter);
void modifyCamera(std::string name, const float *pos, const float *focus, const float *up, float fov);
void addFilm(std::string name, int w, int h, std::string path);
void modifyFilm(std::string name, int w, int h, std::string path);
void addRenderer(std::string name, int adapter, int schedule, std::string const &Camera = "Camera", std::string const &Film = "Film",
                 bool volume = false);
void modifyRenderer(std::string name, int adapter, int schedule, std::string const &Camera = "Camera",
                    std::string const &Film = "Film");
void render(std::string name);
void gvtsync();
void writeimage(std::string name, std::string output = "");

// ---- additions of this engine (not in the reference API) ------------------------------------------------
// shading/routing mode bits of include/gvtb200.h (GVTB200_FLAG_*), RNG seed for area lights and bounces
void setEngineFlags(unsigned flags, unsigned seed = 0);
// the RGBA float image of the last render (what gvtRenderer's myimage->colorbf() returns, IceTComposite.h:106)
const std::vector<float> &framebuffer(int *width = nullptr, int *height = nullptr);
// drop every node (the reference's database lives for the process; tests need a fresh one)
void gvtReset();
// the ObjReader path (data/reader/ObjReader.cpp:192-194): 0-based faces pushed without addFace's filter
void addMeshTriangles0(const std::string name, const unsigned &n, const unsigned *triangles0);
// how many GPUs of this process api::render uses (default 1).  With n > 1 the renderer's schedule type selects the multi-GPU
// tracer, as Renderer.cpp:71-96 selects Tracer<ImageScheduler> / Tracer<DomainScheduler>: Image / AsyncImage = film rows
// dealt over the GPUs with the geometry replicated, Domain = instance k on GPU k % n with exchange rounds over NVLink,
// AsyncDomain = the persistent scheduler.  One host thread per GPU; the image comes back over n PCIe links.
void setGpus(int n);
// handle of the device-side scene behind a renderer (for the multi-GPU drivers), built on first use
void *engineScene(std::string renderer);

} // namespace api
#endif
