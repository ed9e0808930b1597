/* gvtb200.h -- C ABI of the B200-native GraviT render adapter (libgvtb200.so).
 *
 * This is the drop-in boundary for the ray-trace hot path of TACC/GraviT.  Every entry point names the
 * reference interface it replaces (paths relative to /root/reference/src/gvt/render/).  Plain pointers and
 * sizes only; no C++ or torch types.  All pointers are HOST pointers unless the name starts with d_.
 *
 * Conventions
 *   - every function returns 0 on success or a negative gvtb200_status; gvtb200_last_error() gives the text.
 *     Nothing calls exit() (the reference's adapters exit(1)/GVT_ASSERT: EmbreeMeshAdapter.cpp:122,128).
 *   - there is NO CPU fallback: if no CUDA device is usable gvtb200_init fails with GVTB200_ERR_NO_DEVICE.
 *   - one in-flight call per handle; distinct handles may be used from distinct host threads.
 */
#ifndef GVTB200_H
#define GVTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVTB200_API __attribute__((visibility("default")))

typedef enum {
  GVTB200_OK = 0,
  GVTB200_ERR_NO_DEVICE = -1,
  GVTB200_ERR_CUDA = -2,
  GVTB200_ERR_INVALID = -3,
  GVTB200_ERR_CAPACITY = -4,
  GVTB200_ERR_OOM = -5
} gvtb200_status;

/* gvt::render::actor::Ray (actor/Ray.h:45-187): 64-byte payload `rats` (:68-79) inside a 16-aligned 68-byte
 * union (:95) => sizeof(Ray) == 80.  RayVector (:189) is a contiguous array of these. */
typedef struct gvtb200_ray {
  float origin[3];
  float t_min;
  float direction[3];
  float t_max;
  float color[3];
  float t;
  int32_t id;    /* framebuffer index j*W+i */
  int32_t depth; /* remaining bounces */
  float w;
  int32_t type; /* GVTB200_RAY_* == Ray::RayType (actor/Ray.h:50-55) */
  uint8_t pad[16];
} gvtb200_ray;

enum { GVTB200_RAY_PRIMARY = 0, GVTB200_RAY_SHADOW = 1, GVTB200_RAY_SECONDARY = 2 };

/* gvt::render::data::primitives::Material (data/primitives/Material.h:59-89), LAMBERT/PHONG/BLINN fields */
typedef struct gvtb200_material {
  int32_t type; /* 0 LAMBERT, 1 PHONG, 2 BLINN (Material.h:50-57) */
  float kd[3];
  float ks[3];
  float alpha;
} gvtb200_material;

/* gvt::render::data::scene::{PointLight,AmbientLight,AreaLight} (data/scene/Light.h) */
typedef struct gvtb200_light {
  int32_t type; /* 0 point, 1 ambient, 2 area */
  float position[3];
  float color[3];
  float normal[3];
  float width, height;
} gvtb200_light;

enum {
  GVTB200_FLAG_SMOOTH = 1,       /* interpolated vertex normals (the goldens); default = HEAD's FLAT_SHADING
                                    (adapter/embree/EmbreeMeshAdapter.cpp:75,507-524) */
  GVTB200_FLAG_ENTER_INSIDE = 2, /* documented deviation F6: camera rays whose origin is inside an instance box
                                    enter it at t = 0 (reference drops them: actor/RayPacket.h:194) */
  GVTB200_FLAG_SHADOW_TO_LIGHT = 8, /* documented deviation: shadow rays carry t_max = |light - origin| (what
                                    EmbreeMeshAdapter.cpp:347 meant; glm::vec3::length() made it 3) and occlusion is tested up
                                    to the light instead of to infinity (:279).  Off = reference behaviour, under which a
                                    closed interior such as Sibenik is entirely in shadow. */
};

enum { GVTB200_SCHEDULE_IMAGE = 0, GVTB200_SCHEDULE_DOMAIN = 1 }; /* scheduler::ScheduleType (Types.h:48-62) */

typedef struct gvtb200_mesh_s *gvtb200_mesh;
typedef struct gvtb200_scene_s *gvtb200_scene;

/* ---- library ------------------------------------------------------------------------------------------ */
/* replaces api::gvtInit (api/api.cpp:76-98) for the device side: selects the CUDA device, creates streams. */
GVTB200_API int gvtb200_init(int device);
GVTB200_API const char *gvtb200_last_error(void);
/* the device gvtb200_init selected, or -1: lets a second client in the process (the api:: host layer under torchrun,
 * where each rank has already picked its GPU) reuse the selection instead of re-initialising on device 0 */
GVTB200_API int gvtb200_current_device(void);
GVTB200_API int gvtb200_device_info(int *sm_count, size_t *hbm_bytes, int *cc_major, int *cc_minor);
/* scheduling knobs of the trace kernel (results do not depend on them): refill idle lanes when >= refill_at lanes wait,
 * leave the lock-step traversal loop when >= exit_at lanes are not traversing, postpone triangle tests while fewer than
 * postpone_num/8 of the traversing lanes hold a triangle, let a lane visit up to node_burst nodes between two warp votes.
 * Replaces the TBB grain size of EmbreeMeshAdapter.cpp:648. */
GVTB200_API int gvtb200_set_tuning(int refill_at, int exit_at, int postpone_num, int node_burst);
/* tail prefetch: a warp of the trace kernels with at most max_lanes traversing rays (the end of a frame: a few long rays,
 * each a chain of dependent node fetches) requests the sibling nodes it pushes onto its stack into L2 (0 = off; also
 * GVTB200_TAIL_PREFETCH).  Results do not depend on it. */
GVTB200_API int gvtb200_set_tail_prefetch(int max_lanes);
/* tail splitting (persistent kernel): a warp with at most max_lanes traversing rays shares their traversal stacks out among
 * its free lanes, which traverse the donated subtrees as helpers and hand their hits back (the frame's last rays are chains
 * of hundreds of dependent node visits; the chain shrinks with the lanes working on it).  0 = off, default 16; also
 * GVTB200_TAIL_SPLIT.  Results do not depend on it (closest hit: smaller t, then smaller primitive id). */
GVTB200_API int gvtb200_set_tail_split(int max_lanes);
/* how long a kernel of the NVLink peer path waits for another GPU before it gives up and flags the frame (default 4000 ms;
 * also GVTB200_PEER_TIMEOUT_MS).  The reference blocks in MPI_Wait for ever (DomainTracer.h:480-486). */
GVTB200_API int gvtb200_set_peer_timeout_ms(int ms);

/* ---- B1: the engine adapter (Adapter.h:44-88) --------------------------------------------------------- */
/* replaces EmbreeMeshAdapter::EmbreeMeshAdapter (adapter/embree/EmbreeMeshAdapter.cpp:125-162): takes the
 * Mesh the host already holds (vertices, 0-based faces AFTER Mesh::addFace's filter, Mesh.cpp:102-112; vertex
 * normals from Mesh::generateNormals, :116-154, or NULL for flat-only) and builds the BLAS on the GPU
 * (Morton codes -> radix sort -> LBVH -> 8-wide compressed BVH).  face_mat (index into mats, or -1) and
 * vcolors (per vertex) may be NULL (EmbreeMeshAdapter.cpp:536-569 picks vertex colour > face material > mesh). */
GVTB200_API int gvtb200_mesh_create(const float *vertices, uint32_t n_vertices, const uint32_t *faces0, uint32_t n_faces,
                                    const float *vertex_normals, const gvtb200_material *mesh_material,
                                    const gvtb200_material *mats, uint32_t n_mats, const int32_t *face_mat,
                                    const float *vcolors, gvtb200_mesh *out);
GVTB200_API int gvtb200_mesh_destroy(gvtb200_mesh mesh);
/* parity hooks for the BLAS build: Morton key per face (input order), the sorted face order, build stats
 * {n_wide_nodes, n_tris, max_depth, build_us} */
GVTB200_API int gvtb200_mesh_morton(gvtb200_mesh mesh, uint32_t *codes, uint32_t *sorted_faces);
GVTB200_API int gvtb200_mesh_stats(gvtb200_mesh mesh, uint64_t stats[4]);

/* replaces Adapter::trace (Adapter.h:82-84) == EmbreeMeshAdapter::trace (EmbreeMeshAdapter.cpp:625-660):
 * rays[begin,end) are traced against ONE instance of the mesh (m/minv column-major mat4, normi column-major
 * mat3, as glm::value_ptr gives them), mutated in place (bounces), and the outgoing rays -- non-shadow rays
 * that missed, and unoccluded shadow rays carrying their shaded colour -- are APPENDED to moved[*n_moved ..].
 * end == 0 means n_rays (EmbreeMeshAdapter.cpp:642).  GVTB200_ERR_CAPACITY if moved_cap is too small. */
GVTB200_API int gvtb200_trace(gvtb200_mesh mesh, gvtb200_ray *rays, size_t n_rays, size_t begin, size_t end,
                              gvtb200_ray *moved, size_t moved_cap, size_t *n_moved, const float m[16],
                              const float minv[16], const float normi[9], const gvtb200_light *lights, int n_lights,
                              uint32_t flags, uint32_t seed);

/* closest-hit / any-hit queries in mesh-local space (what rtcIntersect8 / rtcOccluded8 return,
 * EmbreeMeshAdapter.cpp:474,375): prim = -1 on miss.  Test hook for hit-record parity. */
typedef struct gvtb200_hit {
  int32_t prim;
  float t, u, v;
} gvtb200_hit;
GVTB200_API int gvtb200_intersect(gvtb200_mesh mesh, const gvtb200_ray *rays, size_t n, gvtb200_hit *hits);
GVTB200_API int gvtb200_occluded(gvtb200_mesh mesh, const gvtb200_ray *rays, size_t n, int32_t *occluded);

/* ---- B2: scene / frame (api/api.h:26-263, Renderer.cpp:37-115, algorithm/ *.h) -------------------------- */
GVTB200_API int gvtb200_scene_create(gvtb200_scene *out);
GVTB200_API int gvtb200_scene_destroy(gvtb200_scene scene);
/* api::addInstance (api/api.cpp:292-322): the host passes what it stored in the context DB -- mat, matinv,
 * normi and the instance box (two transformed corners, :309-312).  Returns the instance id (creation order,
 * which is what BVH::BVH assigns, data/accel/BVH.cpp:48-51). */
GVTB200_API int gvtb200_scene_add_instance(gvtb200_scene scene, gvtb200_mesh mesh, const float m[16], const float minv[16],
                                           const float normi[9], const float box_min[3], const float box_max[3]);
/* api::addPointLight / addAreaLight (api/api.cpp:329-372) */
GVTB200_API int gvtb200_scene_set_lights(gvtb200_scene scene, const gvtb200_light *lights, int n_lights);
/* api::addCamera (api/api.cpp:418-440) + gvtCameraBase::buildTransform (data/scene/gvtCamera.cpp:89-171) */
GVTB200_API int gvtb200_scene_set_camera(gvtb200_scene scene, const float eye[3], const float focus[3], const float up[3],
                                         float fov, int depth, int samples, float jitter);
/* api::addFilm (api/api.cpp:457-468) */
GVTB200_API int gvtb200_scene_set_film(gvtb200_scene scene, int width, int height);
GVTB200_API int gvtb200_scene_set_flags(gvtb200_scene scene, uint32_t flags, uint32_t seed);
/* memory budget of the frame's ray queues, in rays per queue (0 = sized for the whole frame, the default).  The reference
 * grows its std::vector queues without bound (TracerBase.h:403-409); here a frame that needs more than the budget fails with
 * GVTB200_ERR_CAPACITY -- no ray is ever dropped silently. */
GVTB200_API int gvtb200_scene_set_queue_limit(gvtb200_scene scene, size_t max_rays);
/* AbstractTrace::Initialize (algorithm/TracerBase.h:247-308): builds the top-level instance BVH exactly as
 * data/accel/BVH.cpp:44-216 does (same split search, same std::partition) and uploads it. */
GVTB200_API int gvtb200_scene_commit(gvtb200_scene scene);
/* parity hooks: instanceSetID order after the build (BVH.cpp:60-66) */
GVTB200_API int gvtb200_scene_tlas_order(gvtb200_scene scene, int32_t *sorted_ids, int *n);

/* gvtPerspectiveCamera::generateRays (data/scene/gvtCamera.cpp:233-312): all W*H*samples^2 rays, ray-index order */
GVTB200_API int gvtb200_camera_rays(gvtb200_scene scene, gvtb200_ray *out, size_t cap);
/* BVH::intersect<8> (data/accel/BVH.h:61-135): next instance + box-entry t per ray; `from` is excluded */
GVTB200_API int gvtb200_route(gvtb200_scene scene, const gvtb200_ray *rays, size_t n, int from, int32_t *next, float *t);
/* AbstractTrace::shuffleRays (algorithm/TracerBase.h:325-414, non-volume branch): routes `rays`, advances the
 * origin of rays that enter an instance (:394) and returns them grouped by instance in out (out_inst[i] = next
 * instance of out[i]); finished shadow rays are accumulated into the scene framebuffer (:396-399). */
GVTB200_API int gvtb200_shuffle(gvtb200_scene scene, const gvtb200_ray *rays, size_t n, int from, gvtb200_ray *out,
                                int32_t *out_inst, size_t cap, size_t *n_out);
/* IceTComposite::reset / colorbf (composite/IceTComposite.cpp:79-82, IceTComposite.h:106) */
GVTB200_API int gvtb200_framebuffer_clear(gvtb200_scene scene);
GVTB200_API int gvtb200_framebuffer_read(gvtb200_scene scene, float *rgba, size_t n_floats);

/* Tracer<ImageScheduler>::operator() (algorithm/ImageTracer.h:127-269) + gvtRenderer::render (Renderer.cpp:106-115)
 * for pixel rows [row0,row1) -- the whole frame stays on the device: camera rays, routing, two-level traversal,
 * shading, shadow rays, bounces, framebuffer.  rgba_out (W*H*4 floats) may be NULL to leave the result on the
 * device.  counters (may be NULL) receives gvtb200_counters. */
typedef struct gvtb200_counters {
  uint64_t camera_rays;      /* generated */
  uint64_t routed_initial;   /* camera rays that entered an instance */
  uint64_t primary_hits;     /* primary + secondary closest hits shaded */
  uint64_t shadow_generated; /* shadow rays traced */
  uint64_t shadow_unoccluded;
  uint64_t secondary_generated;
  uint64_t fb_adds;
  uint64_t instance_visits; /* (ray, instance) pairs traced == sum of the reference's queue sizes */
  uint64_t nodes_visited;   /* only when profiling counters are enabled */
  uint64_t tris_tested;
  uint64_t rays_exported; /* domain scheduler: rays handed to other ranks */
  uint64_t stack_overflows; /* rays whose traversal stack overflowed; non-zero => GVTB200_ERR_CAPACITY */
  uint64_t queue_overflows; /* rays that found a ray queue (local, or a peer's inbound queue) full; non-zero => the frame calls
                               return GVTB200_ERR_CAPACITY */
} gvtb200_counters;
GVTB200_API int gvtb200_render_frame(gvtb200_scene scene, int row0, int row1, float *rgba_out, gvtb200_counters *counters);
/* same, asynchronous on the scene's stream, nothing copied back (bench / multi-GPU drivers) */
GVTB200_API int gvtb200_render_frame_async(gvtb200_scene scene, int row0, int row1);
/* same, for one shard of the film rows (ImageScheduler, algorithm/ImageTracer.h:111-125 slices the ray array per rank):
 * rows row0 + (b*tile_stride + tile_index)*tile_rows + k, k < tile_rows, b = 0,1,..., clipped to [row0,row1).
 * tile_stride == 1 is the reference's contiguous slice; interleaved tiles balance the load across GPUs. */
GVTB200_API int gvtb200_render_frame_tiles(gvtb200_scene scene, int row0, int row1, int tile_rows, int tile_index, int tile_stride);
GVTB200_API int gvtb200_scene_sync(gvtb200_scene scene);
/* enqueue the scene's work on a caller-owned CUDA stream (the one NCCL collectives use); NULL = the scene's own */
GVTB200_API int gvtb200_scene_set_stream(gvtb200_scene scene, void *cuda_stream);
/* number of kernels this library has launched so far in this process (measurement hook) */
GVTB200_API unsigned long long gvtb200_launch_count(void);
GVTB200_API int gvtb200_scene_counters(gvtb200_scene scene, gvtb200_counters *out);
GVTB200_API int gvtb200_scene_profile_counters(gvtb200_scene scene, int enable);
/* per-launch instrumentation of the frame calls: when enabled every kernel launch of a frame is bracketed by CUDA events on
 * the scene's stream and followed by a snapshot of the (cumulative) counters.  launch_records returns, for the last frame,
 * kind[i] (0 camera+route, 1 closest-hit trace, 2 shade, 3 shadow trace), ms[i] and after[i]; any of the three may be NULL. */
GVTB200_API int gvtb200_scene_profile_launches(gvtb200_scene scene, int enable);
GVTB200_API int gvtb200_scene_launch_records(gvtb200_scene scene, int cap, int *n, int32_t *kind, float *ms, gvtb200_counters *after);
/* device pointer of the RGBA float framebuffer (W*H*4) and the CUDA stream the scene launches on, so that
 * the multi-GPU drivers can hand them to NCCL (replaces icetCompositeImage, composite/IceTComposite.cpp:96) */
GVTB200_API int gvtb200_scene_device_framebuffer(gvtb200_scene scene, void **d_rgba, void **cuda_stream);
/* Parallel read-back of a multi-GPU frame (replaces the gather behind icetCompositeImage, composite/IceTComposite.cpp:96-101,
 * and colorbf(), IceTComposite.h:106): this rank's row tiles (rows (b*tile_stride + tile_index)*tile_rows + k) of the final
 * image -- taken from its own framebuffer (src_rank < 0: ImageScheduler, the rows it rendered) or from attached rank
 * src_rank's framebuffer over NVLink (DomainScheduler: the root holds the composited image) -- are clamped and copied to the
 * SAME rows of rgba_image, a full-size W*H*4 float host image (e.g. a pinned mapping shared by the ranks), over this rank's
 * own PCIe link.  Asynchronous on the scene's stream. */
GVTB200_API int gvtb200_framebuffer_read_rows(gvtb200_scene scene, float *rgba_image, int tile_rows, int tile_index, int tile_stride, int src_rank);
/* final clamp after a cross-GPU sum (localAdd's clamp, composite/IceTComposite.cpp:113-114) */
GVTB200_API int gvtb200_framebuffer_clamp(gvtb200_scene scene);
/* elapsed GPU time (ms) of the last render_frame[_async] on this scene, measured with CUDA events on the
 * scene's stream; kernel_ms[3] = {camera+route, closest-hit trace, shadow trace} of the last wave */
GVTB200_API int gvtb200_scene_last_timing(gvtb200_scene scene, float *frame_ms, float kernel_ms[3]);

/* ---- Tracer<DomainScheduler> (algorithm/DomainTracer.h:115-496) ---------------------------------------- */
/* mpiInstanceMap (DomainTracer.h:115-144): owner rank of every instance; meshes of instances this rank does not
 * own may be NULL in gvtb200_scene_add_instance. */
GVTB200_API int gvtb200_domain_set_owner(gvtb200_scene scene, const int32_t *owner, int n_instances, int rank, int world_size);
/* FilterRaysLocally/shuffleDropRays (:148-181) then the local drain loop (:222-324) until every local queue is
 * empty.  Rays bound for instances owned by other ranks are parked in per-peer export queues. */
GVTB200_API int gvtb200_domain_begin_frame(gvtb200_scene scene);
GVTB200_API int gvtb200_domain_drain_local(gvtb200_scene scene);
/* SendRays phase 1 (:409-415): counts[p] = rays parked for rank p (world_size entries).  phase 2 (:433-463): the
 * parked rays packed peer-major in one DEVICE buffer of 80-byte wire records -- the reference's Ray record
 * (:444-451) whose first pad word carries the target instance id.  *d_export points at it, *n_total = sum(counts).
 * The export queue is emptied. */
GVTB200_API int gvtb200_domain_export(gvtb200_scene scene, uint32_t *counts, void **d_export, size_t *n_total);
/* unpack n_rays received wire records (device pointer) into the local queues */
GVTB200_API int gvtb200_domain_import(gvtb200_scene scene, const void *d_rays, size_t n_rays);
/* rays still queued locally (termination test, DomainTracer.h:339-349) */
GVTB200_API int gvtb200_domain_pending(gvtb200_scene scene, uint64_t *n);

/* ---- NVLink peer path: one process per GPU, peers' memory mapped with CUDA IPC ----------------------------------
 * Replaces, with direct loads/stores/atomics over NVLink 5 / NVSwitch instead of a library collective:
 *   - SendRays' two MPI phases (DomainTracer.h:409-415 counts, :433-463 payload): the trace kernel itself reserves a slot in
 *     the owning rank's inbound ray queue (one remote atomic per warp and owner) and stores the SoA ray there, so the
 *     exchange overlaps the traversal ray by ray and there is no pack / all-to-all / unpack step;
 *   - icetCompositeImage (composite/IceTComposite.cpp:96): ImageScheduler ranks store their clamped film rows into the
 *     root's framebuffer, DomainScheduler ranks add the pixels they touched to it (device-side gather / accumulate).
 * What stays a collective is the termination test (DomainTracer.h:339-349), one 8-byte all-reduce per round.
 *
 * gvtb200_p2p_export allocates this rank's shared objects (framebuffer, two inbound ray queues, a mailbox of flags) and
 * describes them; the ranks exchange the descriptors (any transport: torch.distributed all_gather, MPI, a file) and each
 * calls gvtb200_p2p_attach with all `world` of them.  Descriptors created by the same process (several scenes on one GPU:
 * the single-process tests) are attached by pointer, others through cudaIpcOpenMemHandle. */
typedef struct gvtb200_peer_desc {
  uint8_t fb_handle[64];       /* cudaIpcMemHandle_t of the RGBA float framebuffer */
  uint8_t inbox_handle[2][64]; /* inbound ray queues, used alternately by exchange rounds */
  uint8_t mail_handle[64];     /* flags: go, done[world], error */
  uint64_t raw[4];             /* the same four device pointers, valid inside the creating process */
  uint64_t inbox_cap;          /* rays per inbound queue */
  uint64_t inbox_cap1;         /* the second queue's (only the round-synchronous DomainScheduler uses it) */
  uint64_t fb_pix;
  int32_t pid, device;
} gvtb200_peer_desc;
GVTB200_API int gvtb200_p2p_export(gvtb200_scene scene, int world_size, gvtb200_peer_desc *out);
GVTB200_API int gvtb200_p2p_attach(gvtb200_scene scene, int rank, int world_size, const gvtb200_peer_desc *all);
GVTB200_API int gvtb200_p2p_detach(gvtb200_scene scene);
/* Tracer<ImageScheduler> on `world_size` GPUs: this rank's row tiles (as gvtb200_render_frame_tiles, index = rank, stride =
 * world_size), then the clamped rows go straight into rank `root`'s framebuffer over NVLink.  The root tells the peers when
 * its framebuffer may be overwritten (after its own clear) and waits for their rows before the call's work completes on
 * its stream; a peer that waits longer than the peer time-out gives up and the next gvtb200_scene_sync reports it. */
enum { GVTB200_P2P_CLEAR = 1, GVTB200_P2P_TRACE = 2, GVTB200_P2P_WAIT = 4, GVTB200_P2P_ALL = 7 };
/* phases: GVTB200_P2P_ALL for a frame.  Several scenes of ONE process on ONE GPU (the single-process tests) split the
 * root's call -- CLEAR before the peers' frames, TRACE|WAIT after them -- so that no spinning kernel waits for a launch
 * the host has not made yet. */
GVTB200_API int gvtb200_image_frame_p2p(gvtb200_scene scene, int row0, int row1, int tile_rows, int root, int phases);
/* Tracer<DomainScheduler>, one exchange round: trace what is queued locally (round 0: the camera rays this rank kept in
 * gvtb200_domain_begin_frame; round r > 0: what the peers stored into this rank's inbound queue during round r-1); rays
 * routed to another rank's instance are stored into that rank's inbound queue by the trace kernel.  *d_exported_total
 * (may be NULL) receives the DEVICE address of the frame's cumulative exported-ray counter (uint64) for the termination
 * all-reduce: the frame is over when the sum over ranks did not grow during a round. */
GVTB200_API int gvtb200_domain_round_p2p(gvtb200_scene scene, int round, void **d_exported_total);
/* Persistent DomainScheduler: Tracer<DomainScheduler>::operator() (algorithm/DomainTracer.h:185-368), SendRays (:370-496), the
 * termination test (:339-349) and the composite (composite/IceTComposite.cpp:84-117) as ONE kernel launch per rank and frame;
 * like the reference's asynchronous tracer (scheduler::AsyncDomain, Types.h:53-54; tracer/Domain/DomainTracer.cpp:109-192,
 * votes core/comm/vote/vote.cpp) it has no exchange rounds.  Every rank generates 1/world_size of the camera rays (row tiles of
 * tile_rows rows dealt round-robin), a ray bound for an instance another rank serves is stored into that rank's inbound
 * queue over NVLink by the trace kernel and picked up there by a waiting lane without any launch or host round trip, shading
 * happens in the same kernel (on SMs set aside for it), finished shadow rays are added straight into the root's framebuffer, termination is found on
 * the device (sum of rays retired == sum of rays handed out, read in that order) and the root clamps the image before its
 * kernel ends.  world_size 1 (no peers needed) runs the same kernel on one GPU.  The call is asynchronous on the scene's
 * stream; gvtb200_scene_sync reports time-outs and full queues.  The image is complete in the root's framebuffer. */
GVTB200_API int gvtb200_domain_frame_persist(gvtb200_scene scene, int root, int tile_rows);
/* grid_ctas: CTAs of the persistent kernel (0 = fill the GPU; tests that run several simulated ranks on one GPU bound it so
 * that all of them are resident together); shade_sms: how many SMs shade hits instead of tracing rays (0 = a tenth) */
GVTB200_API int gvtb200_domain_persist_config(gvtb200_scene scene, int grid_ctas, int shade_sms);
/* after a persistent frame: how long this rank was busy (start barrier to its last retired ray, ms), how many rays its
 * inbound queue received and -- for a frame rendered with gvtb200_scene_profile_counters on -- the wide BVH nodes visited per
 * instance on this rank (n_instances entries; input of a replica plan).  Any pointer may be NULL. */
GVTB200_API int gvtb200_domain_persist_stats(gvtb200_scene scene, double *busy_ms, uint64_t *inbox_rays, uint64_t *instance_load, int n_instances);
/* after a persistent frame rendered with gvtb200_scene_profile_counters on: SM cycles summed over this rank's warps --
 * [0] camera phase; tracer warps: [1] waiting for rays, [2] traversal loop, [3] finishing rays (routing, sends, hit records),
 * [4] drawing / polling tickets; shader warps: [5] shading, [6] waiting for hits; [7] warps, [8] shader warps, [9] tracer warps */
GVTB200_API int gvtb200_domain_persist_profile(gvtb200_scene scene, uint64_t prof[16]);
/* the root's view of the last persistent frame, sampled every 16 us by its control warp while it polled for termination:
 * samples[3*k] = microseconds since the start barrier, [3*k+1] = rays retired on all ranks, [3*k+2] = rays handed out on
 * all ranks (up to 256 samples; *n receives the count).  Shows whether a frame is throughput or latency-chain bound. */
GVTB200_API int gvtb200_domain_persist_progress(gvtb200_scene scene, uint32_t *samples, int cap, int *n);
/* Which ranks serve an instance (persistent scheduler only; default: its owner).  ranks[4*i+j], j < 4, are the ranks that
 * hold a BLAS for instance i (-1 = unused), weights[4*i+j] their shares (NULL = equal); a ray bound for the instance picks
 * one by a hash of its direction bits, so the choice is reproducible.  This is how the work of a hot domain is spread:
 * GraviT's hybrid schedulers load a copy of a busy domain on idle ranks (the `Locations` list of a mesh,
 * DomainTracer.h:115-144); here a rank gets its copy from the owner over NVLink (gvtb200_mesh_import).
 * The table belongs to the calling rank and the ranks' tables may differ: a rank that holds a copy of a domain can list
 * itself alone for it ("local first": what meets the domain on this rank is traced on this rank) while ranks that do not
 * hold it list its holders; with a copy of everything everywhere no ray ever leaves its rank. */
GVTB200_API int gvtb200_domain_set_servers(gvtb200_scene scene, const int32_t *ranks, const float *weights, int n_instances);
/* swap the mesh of an instance (NULL = not resident); gvtb200_scene_commit afterwards */
GVTB200_API int gvtb200_scene_set_instance_mesh(gvtb200_scene scene, int instance, gvtb200_mesh mesh);

/* BLAS sharing between ranks: the owner describes its mesh, another rank (process, or device of the same process) imports
 * it without rebuilding anything: GVTB200_IMPORT_COPY = a replica in the importer's HBM, fetched over NVLink;
 * GVTB200_IMPORT_MAP = no copy, the importer traverses the owner's arrays through peer-mapped pointers. */
typedef struct gvtb200_mesh_desc {
  uint8_t handle[6][64]; /* cudaIpcMemHandle_t of: nodes, triangle records, faces, vertex normals, vertex colours, materials */
  uint64_t raw[6];       /* the same device pointers (valid inside the owner's process), 0 = absent */
  uint64_t bytes[6];
  uint64_t tri_slots;
  float mat[8];          /* the mesh material */
  uint32_t n_tris, n_nodes, n_vertices, n_mats;
  int32_t pid, device;
} gvtb200_mesh_desc;
enum { GVTB200_IMPORT_COPY = 0, GVTB200_IMPORT_MAP = 1 };
GVTB200_API int gvtb200_mesh_export(gvtb200_mesh mesh, gvtb200_mesh_desc *out);
GVTB200_API int gvtb200_mesh_import(const gvtb200_mesh_desc *desc, int mode, gvtb200_mesh *out);
/* end of a DomainScheduler frame: every rank adds the pixels it touched to the root's framebuffer; the root waits for all
 * of them, then clamps (IceTComposite::localAdd's clamp, composite/IceTComposite.cpp:113-114). */
GVTB200_API int gvtb200_domain_composite_p2p(gvtb200_scene scene, int root);

#ifdef __cplusplus
}
#endif
#endif /* GVTB200_H */
