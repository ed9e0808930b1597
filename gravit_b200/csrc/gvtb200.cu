// gvtb200.cu -- the C ABI (include/gvtb200.h) over the sm_100a kernels.  Host logic only: device memory, launch
// order, the top-level instance BVH build (a restatement of data/accel/BVH.cpp, it is O(instances) setup work) and
// the camera basis (gvtCamera.cpp:89-171).  There is NO CPU ray-tracing path in this library.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <unistd.h>
#include <string>
#include <vector>

#include "../../include/gvtb200.h"
#include "kernels.cuh"

using namespace gvt;

static_assert(sizeof(gvtb200_ray) == 80, "gvtb200_ray must match gvt::render::actor::Ray (80 bytes)");

namespace {

thread_local std::string t_err;
// One process may drive several GPUs (one host thread per GPU: the C++ multi-GPU tracers of host/api.cpp): gvtb200_init
// selects the device of the CALLING THREAD, every handle remembers the device it was created on, and every call on a
// handle switches to that device first.
#define GVT_MAX_DEVICES 16
struct DevState {
  bool ready = false;
  int sms = 0;
  int trace_bps[2] = { 0, 0 };   // CTAs per SM of k_trace<false/true>
  int persist_bps[2] = { 0, 0 }; // CTAs per SM of k_persist<false/true>
};
DevState g_dev[GVT_MAX_DEVICES];
int g_device = -1;                 // the device of the last successful gvtb200_init of the process
thread_local int t_device = -1;    // the device of this thread's last gvtb200_init
inline int cur_device() { return t_device >= 0 ? t_device : g_device; }
std::atomic<unsigned long long> g_launches{ 0 }; // kernels launched by this library (bench.py reports it)
#define LAUNCHED(n) (g_launches += (n))
unsigned long long g_peer_timeout_ns = 4000000000ull; // gvtb200_set_peer_timeout_ms

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_err = buf;
  return code;
}
#define CU(expr)                                                                                                         \
  do {                                                                                                                   \
    cudaError_t e__ = (expr);                                                                                            \
    if (e__ != cudaSuccess)                                                                                              \
      return fail(e__ == cudaErrorMemoryAllocation ? GVTB200_ERR_OOM : GVTB200_ERR_CUDA, "%s:%d %s: %s", __FILE__,     \
                  __LINE__, #expr, cudaGetErrorString(e__));                                                             \
  } while (0)
#define NEED_INIT()                                                                                                      \
  do {                                                                                                                   \
    if (cur_device() < 0) return fail(GVTB200_ERR_NO_DEVICE, "gvtb200_init has not succeeded: no CUDA device (no CPU fallback)"); \
  } while (0)

inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)std::max<size_t>(1, (n + threads - 1) / threads); }

template <class T> cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }
template <class T> void dfree(T *&p) {
  if (p) cudaFree((void *)p);
  p = nullptr;
}
// BLAS build scratch: stream-ordered allocations from the device's default pool.  gvtb200_init raises the pool's release
// threshold, so after the first build the pool serves them without a driver call (cudaMalloc / cudaFree were 85 % of the
// build time of a 70 K-triangle mesh).
template <class T> cudaError_t salloc(T **p, size_t n, cudaStream_t st) { return cudaMallocAsync((void **)p, std::max<size_t>(n, 1) * sizeof(T), st); }
template <class T> void sfree(T *&p, cudaStream_t st) {
  if (p) cudaFreeAsync((void *)p, st);
  p = nullptr;
}

// ---- SoA ray queue storage --------------------------------------------------------------------------------
struct QueueMem {
  RayQueue q;
  int4 *hit = nullptr; // extra 16-byte column (hit queues only)
  void *slab = nullptr;
  QueueMem() { memset(&q, 0, sizeof(q)); }
  cudaError_t ensure(size_t cap, bool with_hit = false, bool with_ready = false) {
    if (cap < 32) cap = 32;
    if (slab && q.cap >= cap && (!with_hit || hit) && (!with_ready || q.ready)) return cudaSuccess;
    release();
    size_t bytes = cap * ((with_hit ? 84 : 68) + (with_ready ? 4 : 0)) + 256;
    cudaError_t e = cudaMalloc(&slab, bytes);
    if (e != cudaSuccess) {
      slab = nullptr;
      return e;
    }
    char *p = (char *)slab;
    q.count = (unsigned *)p, p += 256;
    q.c0 = (float4 *)p, p += cap * 16;
    q.c1 = (float4 *)p, p += cap * 16;
    q.c2 = (float4 *)p, p += cap * 16;
    q.c3 = (int4 *)p, p += cap * 16;
    if (with_hit) hit = (int4 *)p, p += cap * 16;
    q.inst = (int *)p, p += cap * 4;
    if (with_ready) {
      q.ready = (unsigned *)p;
      e = cudaMemset(q.ready, 0, cap * 4); // frame numbers start at 1
      if (e != cudaSuccess) return e;
    }
    q.cap = (unsigned)cap;
    return cudaMemset(q.count, 0, 256);
  }
  void release() {
    if (slab) cudaFree(slab);
    slab = nullptr;
    hit = nullptr;
    memset(&q, 0, sizeof(q));
  }
};

// Box3D helpers in the reference's operation order (data/primitives/BBox.cpp:33-133)
struct HBox {
  float lo[3], hi[3];
  HBox() {
    for (int i = 0; i < 3; ++i) lo[i] = FLT_MAX, hi[i] = -FLT_MAX;
  }
  void merge(const HBox &o) { // BBox.cpp:55-63
    for (int i = 0; i < 3; ++i) lo[i] = fastmin(o.lo[i], lo[i]), hi[i] = fastmax(o.hi[i], hi[i]);
  }
  int wideRangingBoxDir() const { // BBox.cpp:118-126
    float dx = xsub(hi[0], lo[0]), dy = xsub(hi[1], lo[1]), dz = xsub(hi[2], lo[2]);
    if (dx > dy && dx > dz) return 0;
    if (dy > dz) return 1;
    return 2;
  }
  float centroid(int a) const { return xadd(xmul(lo[a], 0.5f), xmul(hi[a], 0.5f)); } // BBox.cpp:128
  float surfaceArea() const {                                                         // BBox.cpp:130-133
    float dx = xsub(hi[0], lo[0]), dy = xsub(hi[1], lo[1]), dz = xsub(hi[2], lo[2]);
    return xmul(2.f, xadd(xadd(xmul(dx, dy), xmul(dy, dz)), xmul(dz, dx)));
  }
};

} // namespace

// =============================================================================================================
// handles
// =============================================================================================================
struct gvtb200_mesh_s {
  int device = 0;
  unsigned nv = 0, nf = 0, n_mats = 0;
  bool imported = false;           // gvtb200_mesh_import: a replica (own copy) or a mapping of another rank's BLAS
  void *ipc_opened[6] = {};        // cudaIpcOpenMemHandle results kept open (mapped import)
  bool borrowed[6] = {};           // arrays that belong to another mesh of this process (mapped import by pointer)
  float4 *d_nodes = nullptr, *d_tris = nullptr, *d_normals = nullptr, *d_vcolors = nullptr, *d_mats = nullptr;
  int4 *d_faces = nullptr;
  uint32_t *d_keys_unsorted = nullptr, *d_sorted_faces = nullptr;
  Blas blas;
  uint64_t stats[4] = { 0, 0, 0, 0 };
  size_t tri_slots = 0; // 48-byte triangle record slots (GVT_LEAF_MAX per leaf child)
  // adapter-mode (B1) working set, created on first use
  cudaStream_t stream = nullptr;
  Instance *d_inst = nullptr;
  Blas *d_blas = nullptr;
  DevCounters *d_ctr = nullptr;
  unsigned *d_work = nullptr;
  uint4 *d_aos = nullptr;
  size_t aos_cap = 0;
  QueueMem q_in, q_nx, q_sh, q_mv, q_fin, q_hit;
  int4 *d_hits = nullptr;
  size_t hits_cap = 0;
};

namespace {
struct InstHost {
  gvtb200_mesh mesh;
  float m[16], minv[16], normi[9];
  HBox box;
  float centroid[3];
  int owner;
};
struct TNodeHost {
  HBox bbox;
  int left = -1, right = -1, first = 0, count = 0;
};
} // namespace

struct gvtb200_scene_s {
  int device = 0;
  std::vector<InstHost> inst;
  std::vector<gvtb200_light> lights;
  float eye[3] = { 0, 0, 1 }, focus[3] = { 0, 0, 0 }, up[3] = { 0, 1, 0 };
  float fov = 0.785398f;
  int depth = 1, samples = 1, jitter_window = 0;
  int W = 512, H = 512;
  uint32_t flags = 0, seed = 0;
  int rank = 0, world = 1;
  bool committed = false, profile = false;
  size_t queue_limit = 0; // gvtb200_scene_set_queue_limit: rays per queue (0 = sized for the frame)
  // TLAS
  std::vector<TNodeHost> tnodes;
  std::vector<int> work, sorted_ids;
  // device
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
  // per-launch instrumentation (gvtb200_scene_profile_launches): one event pair and one counter snapshot per kernel launch
  enum { MAX_LAUNCH_REC = 64 };
  bool launch_prof = false;
  int n_launch_rec = 0;
  int launch_kind[MAX_LAUNCH_REC];
  cudaEvent_t lev[2 * MAX_LAUNCH_REC] = {};
  DevCounters *d_snap = nullptr;
  Instance *d_inst = nullptr;
  Blas *d_blas = nullptr;
  float4 *d_tnodes = nullptr, *d_tboxes = nullptr;
  Tlas tlas;
  float *d_fb = nullptr;
  size_t fb_pix = 0;
  DevCounters *d_ctr = nullptr;
  unsigned *d_work = nullptr, *d_peer = nullptr;
  QueueMem q_a, q_b, q_sh, q_ex, q_tmp, q_hit;
  uint4 *d_export = nullptr;
  size_t export_cap = 0;
  int *d_next = nullptr;
  float *d_t = nullptr;
  uint4 *d_aos = nullptr;
  size_t aos_cap = 0, route_cap = 0;
  float4 *d_rows = nullptr; // staging of gvtb200_framebuffer_read_rows
  size_t rows_cap = 0;
  float last_frame_ms = 0.f, last_kernel_ms[3] = { 0, 0, 0 };
  bool timing_valid = false;
  // NVLink peer path (gvtb200_p2p_*): this rank's shared objects and the peers' objects mapped with CUDA IPC
  struct {
    bool exported = false, attached = false;
    int rank = 0, world = 1;
    unsigned seq = 0;            // frame sequence number (the same on every rank: all ranks render the same frames)
    QueueMem inbox[2];           // inbound ray queues, alternated by exchange rounds
    Mailbox *d_mail = nullptr;
    RayQueue *d_peer_inbox = nullptr; // [2][world]
    bool private_box = false;         // mailbox + inbox made for a single-rank persistent frame (nothing exported)
    // persistent DomainScheduler (gvtb200_domain_frame_persist)
    QueueMem hitq;                    // closest hits waiting for the shader SMs (ray + hit column + ready words)
    uint4 *d_servers = nullptr;       // per instance: serving ranks + cumulative weights (gvtb200_domain_set_servers)
    std::vector<uint4> servers;
    unsigned long long *d_inst_load = nullptr; // wide nodes visited per instance (counted frames)
    int grid_limit = 0;               // CTAs of the persistent kernel (0 = fill the GPU); the single-GPU tests run several ranks on one device
    int shade_sms = 0;                // SMs that shade instead of tracing (0 = default share)
    float *peer_fb[GVT_P2P_MAX] = {};
    PeerMail mail;
    void *opened[GVT_P2P_MAX][4] = {}; // cudaIpcOpenMemHandle results (closed on detach)
  } p2p;
};

namespace {

// ---- device-wide scan (used by the radix sort) -------------------------------------------------------------
int scan_inplace(unsigned *d, size_t n, std::vector<unsigned *> &levels, size_t level, cudaStream_t st) {
  unsigned tiles = blocks_for(n, RS_TILE);
  if (tiles == 1) {
    k_scan_tile<<<1, RS_THREADS, 0, st>>>(d, (unsigned)n, nullptr);
    return 0;
  }
  if (levels.size() <= level) {
    unsigned *p = nullptr;
    if (salloc(&p, tiles, st) != cudaSuccess) return -1;
    levels.push_back(p);
  }
  unsigned *sums = levels[level];
  k_scan_tile<<<tiles, RS_THREADS, 0, st>>>(d, (unsigned)n, sums);
  if (scan_inplace(sums, tiles, levels, level + 1, st)) return -1;
  k_scan_add<<<tiles, RS_THREADS, 0, st>>>(d, (unsigned)n, sums);
  return 0;
}

TraceTune g_tune = { 16, 20, 2, 2, 0, 16, 24 };

int trace_grid(int dev, bool count) {
  DevState &D = g_dev[dev];
  int idx = count ? 1 : 0;
  if (!D.trace_bps[idx]) {
    int nb = 0;
    if (count)
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<true>, TRACE_THREADS, 0);
    else
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<false>, TRACE_THREADS, 0);
    D.trace_bps[idx] = std::max(1, nb);
    if (const char *e = getenv("GVTB200_TRACE_BLOCKS_PER_SM")) D.trace_bps[idx] = std::max(1, std::min(nb, atoi(e)));
  }
  return D.sms * D.trace_bps[idx];
}
int persist_grid(int dev, bool count) {
  DevState &D = g_dev[dev];
  int idx = count ? 1 : 0;
  if (!D.persist_bps[idx]) {
    int nb = 0;
    if (count)
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_persist<true>, TRACE_THREADS, 0);
    else
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_persist<false>, TRACE_THREADS, 0);
    D.persist_bps[idx] = std::max(1, nb);
    if (const char *e = getenv("GVTB200_PERSIST_BLOCKS_PER_SM")) D.persist_bps[idx] = std::max(1, std::min(nb, atoi(e)));
  }
  return D.sms * D.persist_bps[idx];
}

void launch_trace(int dev, bool count, const FrameParams &fp, const SceneDev &sd, const TraceQueues &tq, int mode, cudaStream_t st) {
  int grid = trace_grid(dev, count);
  LAUNCHED(1);
  if (count)
    k_trace<true><<<grid, TRACE_THREADS, 0, st>>>(fp, sd, tq, mode, g_tune);
  else
    k_trace<false><<<grid, TRACE_THREADS, 0, st>>>(fp, sd, tq, mode, g_tune);
}
void launch_shade(int dev, const FrameParams &fp, const SceneDev &sd, const ShadeQueues &sq, int mode, cudaStream_t st) {
  LAUNCHED(1);
  k_shade<<<g_dev[dev].sms * 8, 256, 0, st>>>(fp, sd, sq, mode);
}

RayQueue null_queue(unsigned *count) {
  RayQueue q;
  memset(&q, 0, sizeof(q));
  q.count = count; // cap 0: every append is counted and discarded
  return q;
}

void pack_instance(Instance &d, const float minv[16], const float normi[9], int mesh, int owner) {
  // column-major inputs: element (row r, col c) = a[4*c + r]
  d.minv0 = make_float4(minv[0], minv[4], minv[8], minv[12]);
  d.minv1 = make_float4(minv[1], minv[5], minv[9], minv[13]);
  d.minv2 = make_float4(minv[2], minv[6], minv[10], minv[14]);
  d.n0 = make_float4(normi[0], normi[3], normi[6], 0.f);
  d.n1 = make_float4(normi[1], normi[4], normi[7], 0.f);
  d.n2 = make_float4(normi[2], normi[5], normi[8], 0.f);
  d.mesh = mesh, d.owner = owner, d.pad0 = d.pad1 = 0;
}

void pack_lights(FrameParams &fp, const gvtb200_light *lights, int n) {
  fp.n_lights = n;
  for (int i = 0; i < n; ++i) {
    Light &L = fp.lights[i];
    L.type = lights[i].type;
    L.pos = v3(lights[i].position[0], lights[i].position[1], lights[i].position[2]);
    L.color = v3(lights[i].color[0], lights[i].color[1], lights[i].color[2]);
    L.normal = v3(lights[i].normal[0], lights[i].normal[1], lights[i].normal[2]);
    L.width = lights[i].width, L.height = lights[i].height;
  }
}

// gvtCameraBase::buildTransform (data/scene/gvtCamera.cpp:89-171, RIGHT_HAND_CAMERA gvtCamera.h:46) and the
// per-frame constants of gvtPerspectiveCamera::generateRays (:233-260)
void build_camera(const gvtb200_scene_s &s, Camera &c) {
  V3 eye = v3(s.eye[0], s.eye[1], s.eye[2]), focus = v3(s.focus[0], s.focus[1], s.focus[2]), up = v3(s.up[0], s.up[1], s.up[2]);
  V3 w = xnormalize(xsub(focus, eye));
  V3 v = xnormalize(up);
  V3 u;
  u.x = xsub(xmul(w.y, v.z), xmul(w.z, v.y));
  u.y = xsub(xmul(w.z, v.x), xmul(w.x, v.z));
  u.z = xsub(xmul(w.x, v.y), xmul(w.y, v.x));
  u = xnormalize(u);
  V3 upv;
  upv.x = xsub(xmul(u.y, w.z), xmul(w.y, u.z));
  upv.y = xsub(xmul(u.z, w.x), xmul(w.z, u.x));
  upv.z = xsub(xmul(u.x, w.y), xmul(w.x, u.y));
  v = xnormalize(upv);
  c.eye = eye, c.u = u, c.v = v, c.w = w;
  float aspect = xdiv((float)s.W, (float)s.H);
  c.vert = tanf((float)((double)s.fov * 0.5));
  c.horz = xmul(tanf((float)((double)s.fov * 0.5)), aspect);
  const float divider = (float)s.samples;
  c.offset = (float)((1.0 / (double)divider) * (double)s.jitter_window);
  c.wmult = xdiv(2.f, (float)(s.W - 1));
  c.hmult = xdiv(2.f, (float)(s.H - 1));
  c.half_sample = xmul((float)s.samples, 0.5f);
  c.contri = xdiv(1.f, (float)(s.samples * s.samples));
  c.W = s.W, c.H = s.H, c.samples = s.samples, c.depth = s.depth;
}

void build_frame_params(const gvtb200_scene_s &s, FrameParams &fp) {
  memset(&fp, 0, sizeof(fp));
  build_camera(s, fp.cam);
  pack_lights(fp, s.lights.data(), (int)s.lights.size());
  fp.flags = s.flags, fp.seed = s.seed, fp.rank = s.rank, fp.world = s.world;
}

SceneDev scene_dev(const gvtb200_scene_s &s) {
  SceneDev sd;
  sd.tlas = s.tlas;
  sd.inst = s.d_inst;
  sd.blas = s.d_blas;
  sd.fb = s.d_fb;
  sd.ctr = s.d_ctr;
  return sd;
}

int upload_rays(const gvtb200_ray *rays, size_t n, uint4 *&d_aos, size_t &aos_cap, QueueMem &q, int inst_or_slot_mode, cudaStream_t st) {
  // inst_or_slot_mode: >= -1 -> constant inst column; -2 -> inst column = slot index
  if (aos_cap < n) {
    dfree(d_aos);
    CU(dalloc(&d_aos, 5 * n));
    aos_cap = n;
  }
  CU(q.ensure(n));
  CU(cudaMemcpyAsync(d_aos, rays, n * sizeof(gvtb200_ray), cudaMemcpyHostToDevice, st));
  if (n) k_aos_to_soa<<<blocks_for(n, 256), 256, 0, st>>>(d_aos, (unsigned)n, q.q, 0u, inst_or_slot_mode);
  unsigned cnt = (unsigned)n;
  CU(cudaMemcpyAsync(q.q.count, &cnt, 4, cudaMemcpyHostToDevice, st));
  return 0;
}

// ---- top-level BVH: data/accel/BVH.cpp:44-216 -------------------------------------------------------------------
float tlas_find_split(gvtb200_scene_s &s, int axis, int start, int end) { // BVH.cpp:173-216
  float minCost = FLT_MAX, splitPoint = 0.f;
  for (int i = start; i < end; ++i) {
    const HBox &ref = s.inst[s.work[i]].box;
    for (int e = 0; e < 2; ++e) {
      float edge = e == 0 ? ref.lo[axis] : ref.hi[axis];
      HBox lb, rb;
      int lc = 0;
      for (int j = start; j < end; ++j) {
        const HBox &bb = s.inst[s.work[j]].box;
        if (bb.centroid(axis) < edge) {
          ++lc;
          lb.merge(bb);
        } else
          rb.merge(bb);
      }
      int rc = end - start - lc;
      // TRAVERSAL_COST is the double literal 0.5 (BVH.cpp:39): the sum is formed in double, then rounded to float
      float la = xmul(lb.surfaceArea(), (float)lc), ra = xmul(rb.surfaceArea(), (float)rc);
      float cost = (float)(0.5 + (double)la + (double)ra);
      if (cost < minCost) minCost = cost, splitPoint = edge;
    }
  }
  return splitPoint;
}
int tlas_build(gvtb200_scene_s &s, int start, int end) { // BVH.cpp:77-171
  int id = (int)s.tnodes.size();
  s.tnodes.push_back(TNodeHost());
  HBox bbox;
  for (int i = start; i < end; ++i) bbox.merge(s.inst[s.work[i]].box);
  int count = end - start;
  auto leaf = [&]() {
    s.tnodes[id].bbox = bbox;
    s.tnodes[id].first = (int)s.sorted_ids.size();
    s.tnodes[id].count = count;
    for (int i = start; i < end; ++i) s.sorted_ids.push_back(s.work[i]);
    return id;
  };
  if (count <= 1) return leaf(); // LEAF_SIZE 1 (:40,95)
  int axis = bbox.wideRangingBoxDir();
  float split = tlas_find_split(s, axis, start, end);
  auto bound = std::partition(s.work.begin() + start, s.work.begin() + end,
                              [&](int k) { return s.inst[k].centroid[axis] < split; }); // :139-140, BVH.h:160-172
  int mid = (int)(bound - s.work.begin());
  if (mid == start || mid == end) return leaf(); // :144
  int l = tlas_build(s, start, mid);
  int r = tlas_build(s, mid, end);
  s.tnodes[id].left = l, s.tnodes[id].right = r, s.tnodes[id].bbox = bbox, s.tnodes[id].count = 0;
  return id;
}
int tlas_depth(const gvtb200_scene_s &s, int n) {
  if (s.tnodes[n].left < 0) return 1;
  return 1 + std::max(tlas_depth(s, s.tnodes[n].left), tlas_depth(s, s.tnodes[n].right));
}

} // namespace

// =============================================================================================================
// library
// =============================================================================================================
extern "C" {

int gvtb200_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    t_device = -1;
    return fail(GVTB200_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= n || device >= GVT_MAX_DEVICES) return fail(GVTB200_ERR_INVALID, "device %d out of range (0..%d)", device, std::min(n, GVT_MAX_DEVICES) - 1);
  CU(cudaSetDevice(device));
  if (g_dev[device].ready) { // another thread (or an earlier call) has set this device up: select it for this thread
    t_device = device;
    g_device = device;
    return GVTB200_OK;
  }
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, device));
  if (p.major < 10) return fail(GVTB200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
  DevState &D = g_dev[device];
  D = DevState();
  D.sms = p.multiProcessorCount;
  if (const char *e = getenv("GVTB200_TUNE")) { // "refill_at,exit_at,postpone_num" (measurement hook)
    int a = 0, b = 0, c = 0, d = 1;
    if (sscanf(e, "%d,%d,%d,%d", &a, &b, &c, &d) >= 3 && gvtb200_set_tuning(a, b, c, d)) return GVTB200_ERR_INVALID;
  }
  if (const char *e = getenv("GVTB200_PEER_TIMEOUT_MS")) gvtb200_set_peer_timeout_ms(atoi(e));
  if (const char *e = getenv("GVTB200_TAIL_PREFETCH")) gvtb200_set_tail_prefetch(atoi(e));
  if (const char *e = getenv("GVTB200_TAIL_SPLIT")) gvtb200_set_tail_split(atoi(e));
  if (const char *e = getenv("GVTB200_PERSIST_TRIPS")) g_tune.persist_trips = std::max(1, atoi(e)); // (measurement hook)
  CU(cudaFree(0));
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { // keep up to 8 GiB of freed build scratch in the pool
    unsigned long long keep = 8ull << 30;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  D.ready = true;
  t_device = device;
  g_device = device;
  return GVTB200_OK;
}

const char *gvtb200_last_error(void) { return t_err.c_str(); }
int gvtb200_current_device(void) { return cur_device(); }
int gvtb200_set_tail_split(int max_lanes) {
  if (max_lanes < 0 || max_lanes > 32) return fail(GVTB200_ERR_INVALID, "lanes %d outside 0..32", max_lanes);
  g_tune.split_below = max_lanes;
  return GVTB200_OK;
}
int gvtb200_set_tail_prefetch(int max_lanes) {
  if (max_lanes < 0 || max_lanes > 32) return fail(GVTB200_ERR_INVALID, "lanes %d outside 0..32", max_lanes);
  g_tune.prefetch_below = max_lanes;
  return GVTB200_OK;
}
int gvtb200_set_peer_timeout_ms(int ms) {
  if (ms < 1) return fail(GVTB200_ERR_INVALID, "timeout %d ms", ms);
  g_peer_timeout_ns = (unsigned long long)ms * 1000000ull;
  return GVTB200_OK;
}

int gvtb200_set_tuning(int refill_at, int exit_at, int postpone_num, int node_burst) {
  if (refill_at < 1 || refill_at > 32 || exit_at < 1 || exit_at > 32 || postpone_num < 0 || postpone_num > 8 || node_burst < 1 || node_burst > 64)
    return fail(GVTB200_ERR_INVALID, "tuning out of range: refill_at %d exit_at %d (1..32), postpone_num %d (0..8), node_burst %d (1..64)",
                refill_at, exit_at, postpone_num, node_burst);
  g_tune.refill_at = refill_at, g_tune.exit_at = exit_at, g_tune.postpone_num = postpone_num, g_tune.node_burst = node_burst;
  return GVTB200_OK;
}

int gvtb200_device_info(int *sm_count, size_t *hbm_bytes, int *cc_major, int *cc_minor) {
  NEED_INIT();
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, cur_device()));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (hbm_bytes) *hbm_bytes = p.totalGlobalMem;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return GVTB200_OK;
}

// =============================================================================================================
// B1: mesh / BLAS
// =============================================================================================================
int gvtb200_mesh_destroy(gvtb200_mesh m) {
  if (!m) return GVTB200_OK;
  cudaSetDevice(m->device);
  { // arrays of an imported mesh may belong to the owner (mapped by pointer or by IPC handle)
    void **arr[6] = { (void **)&m->d_nodes, (void **)&m->d_tris, (void **)&m->d_faces, (void **)&m->d_normals, (void **)&m->d_vcolors, (void **)&m->d_mats };
    for (int k = 0; k < 6; ++k) {
      if (m->ipc_opened[k]) cudaIpcCloseMemHandle(m->ipc_opened[k]), *arr[k] = nullptr;
      if (m->borrowed[k]) *arr[k] = nullptr;
    }
  }
  dfree(m->d_nodes), dfree(m->d_tris), dfree(m->d_normals), dfree(m->d_vcolors), dfree(m->d_mats), dfree(m->d_faces);
  dfree(m->d_keys_unsorted), dfree(m->d_sorted_faces);
  dfree(m->d_inst), dfree(m->d_blas), dfree(m->d_ctr), dfree(m->d_work), dfree(m->d_aos), dfree(m->d_hits);
  m->q_in.release(), m->q_nx.release(), m->q_sh.release(), m->q_mv.release(), m->q_fin.release(), m->q_hit.release();
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
  return GVTB200_OK;
}

static int mesh_build(gvtb200_mesh m, const float *vertices, uint32_t nv, const uint32_t *faces0, uint32_t nf) {
  cudaStream_t st = m->stream;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  float *d_verts = nullptr;
  uint32_t *d_faces3 = nullptr, *d_keys = nullptr, *d_vals = nullptr, *d_keys2 = nullptr, *d_vals2 = nullptr;
  unsigned *d_hist = nullptr, *d_flag = nullptr, *d_counts = nullptr;
  float4 *bn_lo = nullptr, *bn_hi = nullptr, *wnodes = nullptr, *wtris = nullptr, *sah_tb_lo = nullptr, *sah_tb_hi = nullptr;
  uint32_t *sah_idx2 = nullptr;
  int *sah_node_of = nullptr, *sah_node_of2 = nullptr;
  SahNode *sah_act = nullptr, *sah_act2 = nullptr;
  SahBin *sah_bins = nullptr;
  int2 *bn_child = nullptr, *bn_range = nullptr;
  int *bn_parent = nullptr;
  CollapseTask *d_ta = nullptr, *d_tb = nullptr;
  float *dp_cost = nullptr;
  unsigned char *dp_pick = nullptr;
  int4 *ri_pick = nullptr;
  unsigned long long *ri_lock = nullptr;
  unsigned *ri_go = nullptr, *ri_cnt = nullptr, *ri_newpos = nullptr;
  int *ri_lpar = nullptr, *ri_lslot = nullptr;
  std::vector<unsigned *> scan_levels, scan_levels_flag; // per-size scratch of the two device-wide scans (radix histogram, SAH flags)
  int rc = GVTB200_OK;
  auto cleanup = [&]() {
    sfree(d_verts, st), sfree(d_faces3, st), sfree(d_keys, st), sfree(d_vals, st), sfree(d_keys2, st), sfree(d_vals2, st), sfree(d_hist, st), sfree(d_flag, st);
    sfree(sah_tb_lo, st), sfree(sah_tb_hi, st), sfree(sah_idx2, st), sfree(sah_node_of, st), sfree(sah_node_of2, st), sfree(sah_act, st), sfree(sah_act2, st), sfree(sah_bins, st);
    sfree(d_counts, st), sfree(bn_lo, st), sfree(bn_hi, st), sfree(wnodes, st), sfree(wtris, st), sfree(bn_child, st), sfree(bn_range, st), sfree(bn_parent, st), sfree(d_ta, st), sfree(d_tb, st), sfree(dp_cost, st), sfree(dp_pick, st);
    sfree(ri_pick, st), sfree(ri_lock, st), sfree(ri_go, st), sfree(ri_cnt, st), sfree(ri_newpos, st), sfree(ri_lpar, st), sfree(ri_lslot, st);
    for (auto *p : scan_levels) cudaFreeAsync(p, st);
    for (auto *p : scan_levels_flag) cudaFreeAsync(p, st);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
  };
#define CUB(expr)                                                                                                        \
  do {                                                                                                                   \
    cudaError_t e__ = (expr);                                                                                            \
    if (e__ != cudaSuccess) {                                                                                            \
      rc = fail(e__ == cudaErrorMemoryAllocation ? GVTB200_ERR_OOM : GVTB200_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, \
                #expr, cudaGetErrorString(e__));                                                                         \
      cleanup();                                                                                                         \
      return rc;                                                                                                         \
    }                                                                                                                    \
  } while (0)

  const size_t n = nf;
  CUB(salloc(&d_verts, 3 * (size_t)nv, st));
  CUB(salloc(&d_faces3, 3 * n, st));
  CUB(cudaMemcpyAsync(d_verts, vertices, 3 * (size_t)nv * 4, cudaMemcpyHostToDevice, st));
  CUB(cudaMemcpyAsync(d_faces3, faces0, 3 * n * 4, cudaMemcpyHostToDevice, st));
  // mesh bounds in Mesh::computeBoundingBox order (data/primitives/Mesh.cpp:186-199); min/max are exact
  V3 lo = v3(FLT_MAX, FLT_MAX, FLT_MAX), hi = v3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  for (uint32_t i = 0; i < nv; ++i) {
    lo.x = fastmin(lo.x, vertices[3 * i]), lo.y = fastmin(lo.y, vertices[3 * i + 1]), lo.z = fastmin(lo.z, vertices[3 * i + 2]);
    hi.x = fastmax(hi.x, vertices[3 * i]), hi.y = fastmax(hi.y, vertices[3 * i + 1]), hi.z = fastmax(hi.z, vertices[3 * i + 2]);
  }
  BuildCtx c;
  memset(&c, 0, sizeof(c));
  c.verts = d_verts, c.faces = d_faces3, c.n = nf;
  c.mlo = lo, c.mext = xsub(hi, lo);
  c.pad = blas_leaf_pad(c.mext);
  CUB(salloc(&d_keys, n, st)); CUB(salloc(&d_vals, n, st)); CUB(salloc(&d_keys2, n, st)); CUB(salloc(&d_vals2, n, st));
  CUB(dalloc(&m->d_keys_unsorted, n)); CUB(dalloc(&m->d_sorted_faces, n));
  CUB(salloc(&d_counts, 4, st));
  CUB(cudaEventRecord(e0, st));
  unsigned max_depth = 0, n_wnodes = 1, n_wtris = 0;
  if (nf > 0) {
    c.keys = d_keys, c.vals = d_vals;
    k_tri_setup<<<blocks_for(n, 256), 256, 0, st>>>(c);
    CUB(cudaMemcpyAsync(m->d_keys_unsorted, d_keys, n * 4, cudaMemcpyDeviceToDevice, st));
    // LSD radix sort, 4 x 8 bits, stable
    unsigned nblocks = blocks_for(n, RS_TILE);
    CUB(salloc(&d_hist, 256 * (size_t)nblocks, st));
    uint32_t *ka = d_keys, *va = d_vals, *kb = d_keys2, *vb = d_vals2;
    for (int pass = 0; pass < 4; ++pass) {
      k_rs_hist<<<nblocks, RS_THREADS, 0, st>>>(ka, nf, 8 * pass, d_hist, nblocks);
      if (scan_inplace(d_hist, 256 * (size_t)nblocks, scan_levels, 0, st)) CUB(cudaErrorMemoryAllocation);
      k_rs_scatter<<<nblocks, RS_THREADS, 0, st>>>(ka, va, kb, vb, nf, 8 * pass, d_hist, nblocks);
      std::swap(ka, kb), std::swap(va, vb);
    }
    c.keys = ka, c.vals = va; // 4 passes: back in d_keys / d_vals
    CUB(cudaMemcpyAsync(m->d_sorted_faces, va, n * 4, cudaMemcpyDeviceToDevice, st));
    // binary tree: top-down binned SAH over the Morton-ordered triangles (GVTB200_BUILD=lbvh keeps the Karras radix tree)
    CUB(salloc(&bn_lo, 2 * n, st)); CUB(salloc(&bn_hi, 2 * n, st)); CUB(salloc(&bn_child, n, st)); CUB(salloc(&bn_range, n, st));
    CUB(salloc(&bn_parent, 2 * n, st)); CUB(salloc(&d_flag, n + 1, st));
    CUB(cudaMemsetAsync(d_flag, 0, (n + 1) * 4, st));
    CUB(cudaMemsetAsync(bn_parent, 0xFF, 2 * n * 4, st));
    c.bn_lo = bn_lo, c.bn_hi = bn_hi, c.bn_child = bn_child, c.bn_range = bn_range, c.bn_parent = bn_parent, c.bn_flag = d_flag;
    const char *bmode = getenv("GVTB200_BUILD");
    const bool use_sah = !(bmode && !strcmp(bmode, "lbvh")) && nf > 1;
    if (use_sah) {
      SahCtx sc;
      memset(&sc, 0, sizeof(sc));
      const size_t max_big = n / (GVT_SAH_SMALL + 1) + 1;
      CUB(salloc(&sah_tb_lo, n, st)); CUB(salloc(&sah_tb_hi, n, st)); CUB(salloc(&sah_idx2, n, st)); CUB(salloc(&sah_node_of, n, st)); CUB(salloc(&sah_node_of2, n, st));
      CUB(salloc(&sah_act, n / 2 + 1, st)); CUB(salloc(&sah_act2, n / 2 + 1, st)); CUB(salloc(&sah_bins, max_big * 3 * GVT_SAH_BINS, st));
      sc.c = c, sc.tb_lo = sah_tb_lo, sc.tb_hi = sah_tb_hi, sc.idx = va, sc.idx2 = sah_idx2, sc.node_of = sah_node_of, sc.node_of2 = sah_node_of2;
      sc.act = sah_act, sc.act_next = sah_act2, sc.bins = sah_bins, sc.flag = d_flag, sc.counters = d_counts;
      k_sah_tbox<<<blocks_for(n, 256), 256, 0, st>>>(sc);
      CUB(cudaMemsetAsync(sah_node_of, 0, n * 4, st)); // every element belongs to active node 0 (the root)
      SahNode root;
      memset(&root, 0, sizeof(root));
      root.lo = 0, root.hi = (int)nf, root.id = 0, root.bslot = nf > GVT_SAH_SMALL ? 0 : -1;
      CUB(cudaMemcpyAsync(sah_act, &root, sizeof(root), cudaMemcpyHostToDevice, st));
      const int2 root_range = make_int2(0, (int)nf - 1);
      CUB(cudaMemcpyAsync(bn_range, &root_range, sizeof(int2), cudaMemcpyHostToDevice, st));
      unsigned m_act = 1, m_big = nf > GVT_SAH_SMALL ? 1u : 0u, levels = 0;
      while (m_act) {
        if (++levels > 4096) CUB(cudaErrorUnknown);
        unsigned ctr[3] = { 0u, 0u, 0u };
        if (levels == 1) ctr[0] = 1u; // internal id 0 is the root
        // counters[0] persists (next internal id); [1], [2] restart per level
        if (levels == 1) CUB(cudaMemcpyAsync(d_counts, ctr, 12, cudaMemcpyHostToDevice, st));
        else CUB(cudaMemsetAsync(d_counts + 1, 0, 8, st));
        const unsigned n_bins = m_big * 3 * GVT_SAH_BINS;
        k_sah_reset<<<blocks_for(std::max<size_t>(m_act, n_bins), 256), 256, 0, st>>>(sc, m_act, n_bins);
        k_sah_cb<<<blocks_for(n, 256), 256, 0, st>>>(sc);
        if (m_big) k_sah_bin<<<blocks_for(n, 256), 256, 0, st>>>(sc);
        k_sah_split<<<blocks_for(m_act, 64), 64, 0, st>>>(sc, m_act);
        k_sah_flag<<<blocks_for(n + 1, 256), 256, 0, st>>>(sc);
        if (scan_inplace(d_flag, n + 1, scan_levels_flag, 0, st)) CUB(cudaErrorMemoryAllocation);
        k_sah_children<<<blocks_for(m_act, 64), 64, 0, st>>>(sc, m_act);
        k_sah_scatter<<<blocks_for(n, 256), 256, 0, st>>>(sc);
        unsigned out[3];
        CUB(cudaMemcpyAsync(out, d_counts, 12, cudaMemcpyDeviceToHost, st));
        CUB(cudaStreamSynchronize(st));
        m_act = out[1], m_big = out[2];
        std::swap(sc.idx, sc.idx2), std::swap(sc.node_of, sc.node_of2), std::swap(sc.act, sc.act_next);
      }
      // final element order -> c.vals (d_vals), then the leaf boxes
      if (sc.idx != va) CUB(cudaMemcpyAsync(va, sc.idx, n * 4, cudaMemcpyDeviceToDevice, st));
      CUB(cudaMemsetAsync(d_counts, 0, 16, st));
      k_leaf_setup<<<blocks_for(n, 256), 256, 0, st>>>(c);
      // insertion-based optimisation: nodes move to where the surface-area cost grows least (bvh_build.h reinsert_*)
      const int ri_iters = reinsert_iterations(n);
      if (nf >= 4 && ri_iters > 0) {
        const size_t total = 2 * n - 1;
        const unsigned nb = blocks_for(total, 128);
        CUB(salloc(&ri_pick, total, st)); CUB(salloc(&ri_lock, total, st)); CUB(salloc(&ri_go, total, st)); CUB(salloc(&ri_cnt, n, st));
        CUB(salloc(&ri_newpos, n, st)); CUB(salloc(&ri_lpar, n, st)); CUB(salloc(&ri_lslot, n, st));
        c.ri_pick = ri_pick, c.ri_lock = ri_lock, c.ri_go = ri_go, c.ri_cnt = ri_cnt;
        for (int it = 0; it < ri_iters; ++it) {
          CUB(cudaMemsetAsync(ri_lock, 0, total * sizeof(unsigned long long), st));
          k_ri_search<<<nb, 128, 0, st>>>(c);
          k_ri_lock<<<nb, 128, 0, st>>>(c);
          k_ri_check<<<nb, 128, 0, st>>>(c);
          k_ri_confirm<<<nb, 128, 0, st>>>(c);
          k_ri_apply<<<nb, 128, 0, st>>>(c);
          CUB(cudaMemsetAsync(d_flag, 0, n * 4, st));
          k_refit<<<blocks_for(n, 256), 256, 0, st>>>(c);
        }
        // the moves changed the depth-first order of the leaves: renumber them, re-order the elements
        CUB(cudaMemsetAsync(d_flag, 0, n * 4, st));
        k_ri_count<<<blocks_for(n, 256), 256, 0, st>>>(c);
        k_ri_layout_a<<<nb, 128, 0, st>>>(c, ri_newpos, ri_lpar, ri_lslot);
        uint32_t *vals_out = (va == d_vals) ? d_vals2 : d_vals;
        k_ri_layout_b<<<blocks_for(n, 256), 256, 0, st>>>(c, ri_newpos, ri_lpar, ri_lslot, vals_out);
        va = vals_out, c.vals = va;
        k_leaf_setup<<<blocks_for(n, 256), 256, 0, st>>>(c);
      }
    } else {
      k_leaf_setup<<<blocks_for(n, 256), 256, 0, st>>>(c);
      if (nf > 1) {
        CUB(cudaMemsetAsync(d_flag, 0, n * 4, st));
        k_karras<<<blocks_for(n, 256), 256, 0, st>>>(c);
        k_refit<<<blocks_for(n, 256), 256, 0, st>>>(c);
      }
    }
    // optimal widening: which binary nodes become the children of a wide node (GVTB200_WIDEN=greedy keeps the
    // surface-area-greedy rule)
    const char *wmode = getenv("GVTB200_WIDEN");
    if (nf > 1 && !(wmode && !strcmp(wmode, "greedy"))) {
      CUB(salloc(&dp_cost, 7 * n, st));
      CUB(salloc(&dp_pick, 8 * n, st));
      CUB(cudaMemsetAsync(d_flag, 0, n * 4, st));
      c.dp_cost = dp_cost, c.dp_pick = dp_pick;
      k_collapse_dp<<<blocks_for(n, 256), 256, 0, st>>>(c);
    }
    // collapse into 8-wide compressed nodes, one level per launch
    // worst case: one wide node per face, GVT_LEAF_MAX record slots per single-triangle leaf; shrunk below
    CUB(salloc(&wnodes, GVT_NODE_WORDS * n, st));
    CUB(salloc(&wtris, 3 * GVT_LEAF_MAX * n, st));
    CUB(cudaMemsetAsync(wtris, 0, 3 * GVT_LEAF_MAX * n * sizeof(float4), st));
    CUB(salloc(&d_ta, n, st)); CUB(salloc(&d_tb, n, st));
    c.wnodes = wnodes, c.wtris = wtris, c.node_count = d_counts, c.tri_count = d_counts + 1, c.tri_real = d_counts + 3, c.wnode_cap = nf;
    unsigned init[4] = { 1u, 0u, 0u, 0u };
    CUB(cudaMemcpyAsync(d_counts, init, 16, cudaMemcpyHostToDevice, st));
    CollapseTask root;
    root.bnode = 0, root.wnode = 0;
    CUB(cudaMemcpyAsync(d_ta, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    unsigned cur = 1;
    while (cur) {
      ++max_depth;
      CUB(cudaMemsetAsync(d_counts + 2, 0, 4, st));
      k_collapse<<<blocks_for(cur, 128), 128, 0, st>>>(c, d_ta, cur, d_tb, d_counts + 2);
      CUB(cudaMemcpyAsync(&cur, d_counts + 2, 4, cudaMemcpyDeviceToHost, st));
      CUB(cudaStreamSynchronize(st));
      std::swap(d_ta, d_tb);
    }
    unsigned fin[4];
    CUB(cudaMemcpyAsync(fin, d_counts, 16, cudaMemcpyDeviceToHost, st));
    CUB(cudaStreamSynchronize(st));
    n_wnodes = fin[0], n_wtris = fin[3];
    const size_t n_slots = std::max<size_t>(fin[1], 1);
    CUB(dalloc(&m->d_nodes, GVT_NODE_WORDS * (size_t)n_wnodes));
    CUB(cudaMemcpyAsync(m->d_nodes, wnodes, GVT_NODE_WORDS * sizeof(float4) * (size_t)n_wnodes, cudaMemcpyDeviceToDevice, st));
    CUB(dalloc(&m->d_tris, 3 * n_slots));
    CUB(cudaMemcpyAsync(m->d_tris, wtris, 3 * sizeof(float4) * n_slots, cudaMemcpyDeviceToDevice, st));
    m->tri_slots = n_slots;
  } else {
    CUB(dalloc(&m->d_nodes, GVT_NODE_WORDS));
    CUB(dalloc(&m->d_tris, 3));
    n_wnodes = 0;
  }
  CUB(cudaEventRecord(e1, st));
  CUB(cudaStreamSynchronize(st));
  CUB(cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  m->stats[0] = n_wnodes, m->stats[1] = n_wtris, m->stats[2] = max_depth, m->stats[3] = (uint64_t)(ms * 1000.f);
  cleanup();
#undef CUB
  return GVTB200_OK;
}

int gvtb200_mesh_create(const float *vertices, uint32_t n_vertices, const uint32_t *faces0, uint32_t n_faces, const float *vertex_normals,
                        const gvtb200_material *mesh_material, const gvtb200_material *mats, uint32_t n_mats, const int32_t *face_mat,
                        const float *vcolors, gvtb200_mesh *out) {
  NEED_INIT();
  if (!out) return fail(GVTB200_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if ((n_vertices && !vertices) || (n_faces && !faces0)) return fail(GVTB200_ERR_INVALID, "NULL vertex/face array");
  if (n_faces >= 0x7FFFFFFFu / 4) return fail(GVTB200_ERR_INVALID, "too many faces");
  for (size_t i = 0; i < 3 * (size_t)n_faces; ++i)
    if (faces0[i] >= n_vertices) return fail(GVTB200_ERR_INVALID, "face %zu references vertex %u >= %u", i / 3, faces0[i], n_vertices);
  if (face_mat)
    for (uint32_t i = 0; i < n_faces; ++i)
      if (face_mat[i] >= (int32_t)n_mats) return fail(GVTB200_ERR_INVALID, "face %u references material %d >= %u", i, face_mat[i], n_mats);
  CU(cudaSetDevice(cur_device()));
  gvtb200_mesh m = new gvtb200_mesh_s();
  memset(&m->blas, 0, sizeof(Blas));
  m->device = cur_device();
  m->nv = n_vertices, m->nf = n_faces, m->n_mats = (mats && n_mats && face_mat) ? n_mats : 0;
  int rc;
#define CUM(expr)                                                                                                        \
  do {                                                                                                                   \
    cudaError_t e__ = (expr);                                                                                            \
    if (e__ != cudaSuccess) {                                                                                            \
      rc = fail(e__ == cudaErrorMemoryAllocation ? GVTB200_ERR_OOM : GVTB200_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, \
                #expr, cudaGetErrorString(e__));                                                                         \
      gvtb200_mesh_destroy(m);                                                                                           \
      return rc;                                                                                                         \
    }                                                                                                                    \
  } while (0)
  CUM(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  rc = mesh_build(m, vertices, n_vertices, faces0, n_faces);
  if (rc) {
    gvtb200_mesh_destroy(m);
    return rc;
  }
  { // shading tables
    std::vector<int4> f4(std::max<size_t>(1, n_faces));
    for (uint32_t i = 0; i < n_faces; ++i)
      f4[i] = make_int4((int)faces0[3 * i], (int)faces0[3 * i + 1], (int)faces0[3 * i + 2], face_mat ? face_mat[i] : -1);
    CUM(dalloc(&m->d_faces, f4.size()));
    CUM(cudaMemcpy(m->d_faces, f4.data(), f4.size() * sizeof(int4), cudaMemcpyHostToDevice));
    auto up3 = [&](const float *src, float4 *&dst) -> cudaError_t {
      std::vector<float4> t(std::max<size_t>(1, n_vertices));
      for (uint32_t i = 0; i < n_vertices; ++i) t[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
      cudaError_t e = dalloc(&dst, t.size());
      if (e != cudaSuccess) return e;
      return cudaMemcpy(dst, t.data(), t.size() * sizeof(float4), cudaMemcpyHostToDevice);
    };
    if (vertex_normals) CUM(up3(vertex_normals, m->d_normals));
    if (vcolors) CUM(up3(vcolors, m->d_vcolors));
    auto packm = [](const gvtb200_material &a, float4 &p0, float4 &p1) {
      p0 = make_float4(u2f((uint32_t)a.type), a.kd[0], a.kd[1], a.kd[2]);
      p1 = make_float4(a.ks[0], a.ks[1], a.ks[2], a.alpha);
    };
    gvtb200_material def; // Material.h:62-75 defaults
    def.type = 0, def.kd[0] = def.kd[1] = def.kd[2] = .5f, def.ks[0] = def.ks[1] = def.ks[2] = .5f, def.alpha = 1.f;
    packm(mesh_material ? *mesh_material : def, m->blas.mat0, m->blas.mat1);
    if (mats && n_mats && face_mat) {
      std::vector<float4> t(2 * (size_t)n_mats);
      for (uint32_t i = 0; i < n_mats; ++i) packm(mats[i], t[2 * i], t[2 * i + 1]);
      CUM(dalloc(&m->d_mats, t.size()));
      CUM(cudaMemcpy(m->d_mats, t.data(), t.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
  }
#undef CUM
  m->blas.nodes = m->d_nodes, m->blas.tris = m->d_tris, m->blas.faces = m->d_faces, m->blas.normals = m->d_normals;
  m->blas.vcolors = m->d_vcolors, m->blas.mats = m->d_mats;
  m->blas.n_tris = n_faces, m->blas.n_nodes = (unsigned)m->stats[0];
  *out = m;
  return GVTB200_OK;
}

int gvtb200_mesh_morton(gvtb200_mesh m, uint32_t *codes, uint32_t *sorted_faces) {
  NEED_INIT();
  if (!m) return fail(GVTB200_ERR_INVALID, "mesh is NULL");
  if (m->imported) return fail(GVTB200_ERR_INVALID, "an imported mesh holds no build records");
  if (codes && m->nf) CU(cudaMemcpy(codes, m->d_keys_unsorted, (size_t)m->nf * 4, cudaMemcpyDeviceToHost));
  if (sorted_faces && m->nf) CU(cudaMemcpy(sorted_faces, m->d_sorted_faces, (size_t)m->nf * 4, cudaMemcpyDeviceToHost));
  return GVTB200_OK;
}
int gvtb200_mesh_stats(gvtb200_mesh m, uint64_t stats[4]) {
  if (!m) return fail(GVTB200_ERR_INVALID, "mesh is NULL");
  memcpy(stats, m->stats, sizeof(m->stats));
  return GVTB200_OK;
}

static int mesh_adapter_ctx(gvtb200_mesh m) {
  if (!m->d_inst) {
    CU(dalloc(&m->d_inst, 1));
    CU(dalloc(&m->d_blas, 1));
    CU(dalloc(&m->d_ctr, 1));
    CU(dalloc(&m->d_work, 64));
    CU(cudaMemcpy(m->d_blas, &m->blas, sizeof(Blas), cudaMemcpyHostToDevice));
  }
  return 0;
}

static int gvtb200_intersect_impl(gvtb200_mesh m, const gvtb200_ray *rays, size_t n, int any, int4 *hits_host) {
  NEED_INIT();
  if (!m) return fail(GVTB200_ERR_INVALID, "mesh is NULL");
  if (n == 0) return GVTB200_OK;
  if (n > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "too many rays");
  CU(cudaSetDevice(m->device));
  if (int rc = mesh_adapter_ctx(m)) return rc;
  if (int rc = upload_rays(rays, n, m->d_aos, m->aos_cap, m->q_in, 0, m->stream)) return rc;
  if (m->hits_cap < n) {
    dfree(m->d_hits);
    CU(dalloc(&m->d_hits, n));
    m->hits_cap = n;
  }
  k_intersect<<<blocks_for(n, 128), 128, 0, m->stream>>>(m->blas, m->q_in.q, (unsigned)n, any, m->d_hits);
  CU(cudaMemcpyAsync(hits_host, m->d_hits, n * sizeof(int4), cudaMemcpyDeviceToHost, m->stream));
  CU(cudaStreamSynchronize(m->stream));
  CU(cudaGetLastError());
  return GVTB200_OK;
}

int gvtb200_intersect(gvtb200_mesh m, const gvtb200_ray *rays, size_t n, gvtb200_hit *hits) {
  static_assert(sizeof(gvtb200_hit) == sizeof(int4), "hit record is 16 bytes");
  if (n && (!rays || !hits)) return fail(GVTB200_ERR_INVALID, "NULL rays/hits");
  return gvtb200_intersect_impl(m, rays, n, 0, reinterpret_cast<int4 *>(hits));
}
int gvtb200_occluded(gvtb200_mesh m, const gvtb200_ray *rays, size_t n, int32_t *occluded) {
  if (n && (!rays || !occluded)) return fail(GVTB200_ERR_INVALID, "NULL rays/occluded");
  std::vector<int4> h(n);
  int rc = gvtb200_intersect_impl(m, rays, n, 1, h.data());
  if (rc) return rc;
  for (size_t i = 0; i < n; ++i) occluded[i] = h[i].x >= 0 ? 1 : 0;
  return GVTB200_OK;
}

int gvtb200_trace(gvtb200_mesh m, gvtb200_ray *rays, size_t n_rays, size_t begin, size_t end, gvtb200_ray *moved, size_t moved_cap,
                  size_t *n_moved, const float mat[16], const float minv[16], const float normi[9], const gvtb200_light *lights,
                  int n_lights, uint32_t flags, uint32_t seed) {
  NEED_INIT();
  if (!m || !n_moved || !minv || !normi || (n_lights && !lights)) return fail(GVTB200_ERR_INVALID, "NULL argument");
  (void)mat; // the instance matrix itself is not needed: rays are taken to mesh space with minv, normals back with normi
  if (n_lights < 0 || n_lights > GVT_MAX_LIGHTS) return fail(GVTB200_ERR_INVALID, "at most %d lights", GVT_MAX_LIGHTS);
  if (end == 0) end = n_rays; // EmbreeMeshAdapter.cpp:642
  if (begin > end || end > n_rays) return fail(GVTB200_ERR_INVALID, "bad range [%zu,%zu) of %zu", begin, end, n_rays);
  const size_t n = end - begin;
  if (n == 0) return GVTB200_OK;
  if (!rays || (moved_cap && !moved)) return fail(GVTB200_ERR_INVALID, "NULL rays/moved");
  if (n * (size_t)(1 + n_lights) > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "too many rays in one call");
  CU(cudaSetDevice(m->device));
  if (int rc = mesh_adapter_ctx(m)) return rc;
  cudaStream_t st = m->stream;
  int max_depth = 1; // waves to run; a ray's depth is caller input: bounded (Russian roulette ends real paths long before)
  for (size_t i = begin; i < end; ++i) max_depth = std::max(max_depth, std::min(rays[i].depth, 64));
  Instance in;
  pack_instance(in, minv, normi, 0, 0);
  CU(cudaMemcpyAsync(m->d_inst, &in, sizeof(in), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(m->d_ctr, 0, sizeof(DevCounters), st));
  FrameParams fp;
  memset(&fp, 0, sizeof(fp));
  pack_lights(fp, lights, n_lights);
  fp.flags = flags, fp.seed = seed, fp.rank = 0, fp.world = 1;
  SceneDev sd;
  memset(&sd, 0, sizeof(sd));
  sd.inst = m->d_inst, sd.blas = m->d_blas, sd.ctr = m->d_ctr;
  if (int rc = upload_rays(rays + begin, n, m->d_aos, m->aos_cap, m->q_in, -2, st)) return rc;
  CU(m->q_fin.ensure(n));
  k_aos_to_soa<<<blocks_for(n, 256), 256, 0, st>>>(m->d_aos, (unsigned)n, m->q_fin.q, 0u, -2);
  CU(m->q_nx.ensure(n));
  CU(m->q_sh.ensure(n * (size_t)std::max(1, n_lights)));
  // every input ray can be handed back, every hit can emit one unoccluded shadow ray per light per bounce
  const size_t mv_need = n * (size_t)(1 + n_lights * max_depth);
  CU(m->q_mv.ensure(std::min<size_t>(mv_need, std::max<size_t>(moved_cap, n))));
  CU(cudaMemsetAsync(m->q_mv.q.count, 0, 4, st));
  CU(m->q_hit.ensure(n, true));
  TraceQueues tq;
  ShadeQueues sq;
  memset(&tq, 0, sizeof(tq));
  QueueMem *qa = &m->q_in, *qb = &m->q_nx;
  for (int wave = 0; wave < max_depth; ++wave) {
    CU(cudaMemsetAsync(m->d_work, 0, 8, st));
    CU(cudaMemsetAsync(qb->q.count, 0, 4, st));
    CU(cudaMemsetAsync(m->q_sh.q.count, 0, 4, st));
    CU(cudaMemsetAsync(m->q_hit.q.count, 0, 4, st));
    // closest-hit wave: misses -> moved, hits -> hit queue
    tq.in = qa->q, tq.hq = m->q_hit.q, tq.hit = m->q_hit.hit, tq.mvq = m->q_mv.q, tq.exq = null_queue(m->d_work + 2), tq.fin = m->q_fin.q;
    tq.work = m->d_work;
    launch_trace(m->device, false, fp, sd, tq, 1, st);
    // shading: shadow rays -> q_sh, bounces -> qb, final states -> fin
    sq.hq = m->q_hit.q, sq.hit = m->q_hit.hit, sq.shq = m->q_sh.q, sq.nxq = qb->q, sq.fin = m->q_fin.q;
    launch_shade(m->device, fp, sd, sq, 1, st);
    // shadow wave: unoccluded -> moved
    tq.in = m->q_sh.q, tq.hq = null_queue(m->d_work + 3), tq.hit = nullptr, tq.work = m->d_work + 1;
    launch_trace(m->device, false, fp, sd, tq, 1, st);
    std::swap(qa, qb);
  }
  unsigned n_mv = 0;
  CU(cudaMemcpyAsync(&n_mv, m->q_mv.q.count, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  if (n_mv > m->q_mv.q.cap || *n_moved + n_mv > moved_cap)
    return fail(GVTB200_ERR_CAPACITY, "moved_rays capacity %zu too small: %zu + %u outgoing rays", moved_cap, *n_moved, n_mv);
  // rays back (mutated in place) and moved rays appended
  size_t need = std::max<size_t>(n, n_mv);
  if (m->aos_cap < need) {
    dfree(m->d_aos);
    CU(dalloc(&m->d_aos, 5 * need));
    m->aos_cap = need;
  }
  k_soa_to_aos<<<blocks_for(n, 256), 256, 0, st>>>(m->q_fin.q, (unsigned)n, m->d_aos);
  CU(cudaMemcpyAsync(rays + begin, m->d_aos, n * sizeof(gvtb200_ray), cudaMemcpyDeviceToHost, st));
  if (n_mv) {
    CU(cudaStreamSynchronize(st));
    k_soa_to_aos<<<blocks_for(n_mv, 256), 256, 0, st>>>(m->q_mv.q, n_mv, m->d_aos);
    CU(cudaMemcpyAsync(moved + *n_moved, m->d_aos, (size_t)n_mv * sizeof(gvtb200_ray), cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  *n_moved += n_mv;
  return GVTB200_OK;
}

// =============================================================================================================
// B2: scene / frame
// =============================================================================================================
int gvtb200_scene_create(gvtb200_scene *out) {
  NEED_INIT();
  if (!out) return fail(GVTB200_ERR_INVALID, "out is NULL");
  CU(cudaSetDevice(cur_device()));
  gvtb200_scene s = new gvtb200_scene_s();
  s->device = cur_device();
  memset(&s->tlas, 0, sizeof(Tlas));
  cudaError_t e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
  s->stream = s->own_stream;
  for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaEventCreate(&s->ev[i]);
  if (e == cudaSuccess) e = dalloc(&s->d_ctr, 1);
  if (e == cudaSuccess) e = dalloc(&s->d_work, 64);
  if (e == cudaSuccess) e = dalloc(&s->d_peer, 1024);
  if (e == cudaSuccess) e = cudaMemset(s->d_ctr, 0, sizeof(DevCounters));
  if (e != cudaSuccess) {
    gvtb200_scene_destroy(s);
    return fail(GVTB200_ERR_CUDA, "scene_create: %s", cudaGetErrorString(e));
  }
  *out = s;
  return GVTB200_OK;
}
int gvtb200_scene_destroy(gvtb200_scene s) {
  if (!s) return GVTB200_OK;
  if (s->stream) cudaStreamSynchronize(s->stream);
  gvtb200_p2p_detach(s);
  dfree(s->d_inst), dfree(s->d_blas), dfree(s->d_tnodes), dfree(s->d_tboxes), dfree(s->d_fb), dfree(s->d_ctr), dfree(s->d_work), dfree(s->d_peer);
  dfree(s->d_export), dfree(s->d_next), dfree(s->d_t), dfree(s->d_aos), dfree(s->d_rows);
  s->q_a.release(), s->q_b.release(), s->q_sh.release(), s->q_ex.release(), s->q_tmp.release(), s->q_hit.release();
  for (auto &e : s->ev)
    if (e) cudaEventDestroy(e);
  for (auto &e : s->lev)
    if (e) cudaEventDestroy(e);
  dfree(s->d_snap);
  if (s->own_stream) cudaStreamDestroy(s->own_stream);
  delete s;
  return GVTB200_OK;
}
int gvtb200_scene_add_instance(gvtb200_scene s, gvtb200_mesh mesh, const float m[16], const float minv[16], const float normi[9],
                               const float box_min[3], const float box_max[3]) {
  if (!s || !m || !minv || !normi || !box_min || !box_max) return fail(GVTB200_ERR_INVALID, "NULL argument");
  InstHost in;
  in.mesh = mesh;
  memcpy(in.m, m, 64), memcpy(in.minv, minv, 64), memcpy(in.normi, normi, 36);
  for (int i = 0; i < 3; ++i) in.box.lo[i] = box_min[i], in.box.hi[i] = box_max[i];
  for (int i = 0; i < 3; ++i) in.centroid[i] = in.box.centroid(i); // api.cpp:321
  in.owner = 0;
  s->inst.push_back(in);
  s->committed = false;
  return (int)s->inst.size() - 1;
}
int gvtb200_scene_set_lights(gvtb200_scene s, const gvtb200_light *lights, int n) {
  if (!s || (n && !lights)) return fail(GVTB200_ERR_INVALID, "NULL argument");
  if (n < 0 || n > GVT_MAX_LIGHTS) return fail(GVTB200_ERR_INVALID, "at most %d lights", GVT_MAX_LIGHTS);
  s->lights.assign(lights, lights + n);
  return GVTB200_OK;
}
int gvtb200_scene_set_camera(gvtb200_scene s, const float eye[3], const float focus[3], const float up[3], float fov, int depth,
                             int samples, float jitter) {
  if (!s || !eye || !focus || !up) return fail(GVTB200_ERR_INVALID, "NULL argument");
  if (samples < 1 || depth < 0) return fail(GVTB200_ERR_INVALID, "samples must be >= 1 and depth >= 0");
  memcpy(s->eye, eye, 12), memcpy(s->focus, focus, 12), memcpy(s->up, up, 12);
  s->fov = fov, s->depth = depth, s->samples = samples;
  s->jitter_window = (int)jitter; // gvtCamera.h:136 / gvtCamera.cpp:202: an int
  return GVTB200_OK;
}
int gvtb200_scene_set_film(gvtb200_scene s, int w, int h) {
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  if (w < 2 || h < 2 || (size_t)w * h > 0x40000000u) return fail(GVTB200_ERR_INVALID, "film %dx%d not supported", w, h);
  s->W = w, s->H = h;
  return GVTB200_OK;
}
int gvtb200_scene_set_queue_limit(gvtb200_scene s, size_t max_rays) {
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  s->queue_limit = max_rays;
  return GVTB200_OK;
}
int gvtb200_scene_set_flags(gvtb200_scene s, uint32_t flags, uint32_t seed) {
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  s->flags = flags, s->seed = seed;
  return GVTB200_OK;
}

int gvtb200_scene_commit(gvtb200_scene s) {
  NEED_INIT();
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  CU(cudaSetDevice(s->device));
  const int n = (int)s->inst.size();
  s->tnodes.clear(), s->sorted_ids.clear();
  s->work.resize(n);
  for (int i = 0; i < n; ++i) s->work[i] = i; // ids 0..n-1 in creation order (BVH.cpp:48-51)
  if (n) tlas_build(*s, 0, n);
  if (n && tlas_depth(*s, 0) > GVT_TLAS_STACK) return fail(GVTB200_ERR_CAPACITY, "top-level BVH deeper than %d", GVT_TLAS_STACK);
  // flatten
  std::vector<float4> tn(2 * std::max<size_t>(1, s->tnodes.size())), tb(2 * std::max<size_t>(1, (size_t)n));
  for (size_t i = 0; i < s->tnodes.size(); ++i) {
    const TNodeHost &t = s->tnodes[i];
    int a = t.left >= 0 ? t.left : ~t.first, b = t.left >= 0 ? t.right : t.count;
    tn[2 * i] = make_float4(t.bbox.lo[0], t.bbox.lo[1], t.bbox.lo[2], u2f((uint32_t)a));
    tn[2 * i + 1] = make_float4(t.bbox.hi[0], t.bbox.hi[1], t.bbox.hi[2], u2f((uint32_t)b));
  }
  for (int i = 0; i < n; ++i) {
    const InstHost &in = s->inst[s->sorted_ids[i]];
    tb[2 * i] = make_float4(in.box.lo[0], in.box.lo[1], in.box.lo[2], u2f((uint32_t)s->sorted_ids[i]));
    tb[2 * i + 1] = make_float4(in.box.hi[0], in.box.hi[1], in.box.hi[2], 0.f);
  }
  dfree(s->d_tnodes), dfree(s->d_tboxes), dfree(s->d_inst), dfree(s->d_blas);
  CU(dalloc(&s->d_tnodes, tn.size()));
  CU(dalloc(&s->d_tboxes, tb.size()));
  CU(cudaMemcpy(s->d_tnodes, tn.data(), tn.size() * sizeof(float4), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(s->d_tboxes, tb.data(), tb.size() * sizeof(float4), cudaMemcpyHostToDevice));
  s->tlas.nodes = s->d_tnodes, s->tlas.boxes = s->d_tboxes, s->tlas.n_nodes = (int)s->tnodes.size(), s->tlas.n_inst = n;
  // instance + BLAS tables
  std::map<gvtb200_mesh, int> mesh_index;
  std::vector<Blas> blas;
  std::vector<Instance> di(std::max(1, n));
  for (int i = 0; i < n; ++i) {
    const InstHost &in = s->inst[i];
    int mi = -1;
    if (in.mesh) {
      auto it = mesh_index.find(in.mesh);
      if (it == mesh_index.end()) {
        mi = (int)blas.size();
        mesh_index[in.mesh] = mi;
        blas.push_back(in.mesh->blas);
      } else
        mi = it->second;
    } else if (s->world == 1 || in.owner == s->rank)
      return fail(GVTB200_ERR_INVALID, "instance %d has no mesh but is owned by this rank", i);
    pack_instance(di[i], in.minv, in.normi, mi, in.owner);
  }
  if (blas.empty()) {
    Blas b;
    memset(&b, 0, sizeof(b));
    blas.push_back(b);
  }
  CU(dalloc(&s->d_inst, di.size()));
  CU(dalloc(&s->d_blas, blas.size()));
  CU(cudaMemcpy(s->d_inst, di.data(), di.size() * sizeof(Instance), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(s->d_blas, blas.data(), blas.size() * sizeof(Blas), cudaMemcpyHostToDevice));
  s->committed = true;
  return GVTB200_OK;
}

int gvtb200_scene_tlas_order(gvtb200_scene s, int32_t *sorted_ids, int *n) {
  if (!s || !s->committed) return fail(GVTB200_ERR_INVALID, "scene not committed");
  if (sorted_ids) memcpy(sorted_ids, s->sorted_ids.data(), s->sorted_ids.size() * 4);
  if (n) *n = (int)s->tnodes.size();
  return GVTB200_OK;
}

static int scene_fb(gvtb200_scene s) {
  size_t npix = (size_t)s->W * s->H;
  if (s->fb_pix != npix) {
    if (s->p2p.exported) return fail(GVTB200_ERR_INVALID, "film size changed while the framebuffer is shared with peers: gvtb200_p2p_detach first");
    dfree(s->d_fb);
    CU(dalloc(&s->d_fb, npix * 4));
    s->fb_pix = npix;
    CU(cudaMemsetAsync(s->d_fb, 0, npix * 16, s->stream));
  }
  return 0;
}
#define NEED_SCENE()                                                                                                     \
  do {                                                                                                                   \
    NEED_INIT();                                                                                                         \
    if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");                                                          \
    if (!s->committed) return fail(GVTB200_ERR_INVALID, "scene not committed (gvtb200_scene_commit)");                  \
    CU(cudaSetDevice(s->device));                                                                                         \
  } while (0)

int gvtb200_camera_rays(gvtb200_scene s, gvtb200_ray *out, size_t cap) {
  NEED_INIT();
  if (!s || !out) return fail(GVTB200_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(s->device));
  size_t n = (size_t)s->W * s->H * s->samples * s->samples;
  if (cap < n) return fail(GVTB200_ERR_CAPACITY, "need room for %zu rays", n);
  if (n > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "too many rays");
  FrameParams fp;
  build_frame_params(*s, fp);
  CU(s->q_tmp.ensure(n));
  if (s->aos_cap < n) {
    dfree(s->d_aos);
    CU(dalloc(&s->d_aos, 5 * n));
    s->aos_cap = n;
  }
  k_camera_only<<<blocks_for(n, 256), 256, 0, s->stream>>>(fp, n, s->q_tmp.q);
  k_soa_to_aos<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->q_tmp.q, (unsigned)n, s->d_aos);
  CU(cudaMemcpyAsync(out, s->d_aos, n * sizeof(gvtb200_ray), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  return GVTB200_OK;
}

int gvtb200_route(gvtb200_scene s, const gvtb200_ray *rays, size_t n, int from, int32_t *next, float *t) {
  NEED_SCENE();
  if (n == 0) return GVTB200_OK;
  if (!rays || !next || !t) return fail(GVTB200_ERR_INVALID, "NULL argument");
  if (n > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "too many rays");
  if (int rc = upload_rays(rays, n, s->d_aos, s->aos_cap, s->q_tmp, -1, s->stream)) return rc;
  if (s->route_cap < n) {
    dfree(s->d_next), dfree(s->d_t);
    CU(dalloc(&s->d_next, n));
    CU(dalloc(&s->d_t, n));
    s->route_cap = n;
  }
  FrameParams fp;
  build_frame_params(*s, fp);
  k_route<<<blocks_for(n, 128), 128, 0, s->stream>>>(fp, scene_dev(*s), s->q_tmp.q, (unsigned)n, from, s->d_next, s->d_t);
  CU(cudaMemcpyAsync(next, s->d_next, n * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(t, s->d_t, n * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  return GVTB200_OK;
}

int gvtb200_shuffle(gvtb200_scene s, const gvtb200_ray *rays, size_t n, int from, gvtb200_ray *out, int32_t *out_inst, size_t cap,
                    size_t *n_out) {
  NEED_SCENE();
  if (!n_out) return fail(GVTB200_ERR_INVALID, "NULL argument");
  *n_out = 0;
  if (n == 0) return GVTB200_OK;
  if (!rays || !out || !out_inst) return fail(GVTB200_ERR_INVALID, "NULL argument");
  if (n > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "too many rays");
  if (int rc = scene_fb(s)) return rc;
  if (int rc = upload_rays(rays, n, s->d_aos, s->aos_cap, s->q_tmp, -1, s->stream)) return rc;
  CU(s->q_a.ensure(n));
  CU(cudaMemsetAsync(s->q_a.q.count, 0, 4, s->stream));
  FrameParams fp;
  build_frame_params(*s, fp);
  k_shuffle<<<blocks_for(n, 128), 128, 0, s->stream>>>(fp, scene_dev(*s), s->q_tmp.q, (unsigned)n, from, s->q_a.q);
  unsigned cnt = 0;
  CU(cudaMemcpyAsync(&cnt, s->q_a.q.count, 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  if (cnt > cap) return fail(GVTB200_ERR_CAPACITY, "need room for %u rays", cnt);
  if (cnt) {
    std::vector<gvtb200_ray> tmp(cnt);
    std::vector<int> ti(cnt);
    k_soa_to_aos<<<blocks_for(cnt, 256), 256, 0, s->stream>>>(s->q_a.q, cnt, s->d_aos);
    CU(cudaMemcpyAsync(tmp.data(), s->d_aos, (size_t)cnt * sizeof(gvtb200_ray), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(ti.data(), s->q_a.q.inst, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    // group by target instance (the reference keeps one queue per instance, TracerBase.h:403-409)
    std::vector<unsigned> order(cnt);
    for (unsigned i = 0; i < cnt; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return ti[a] < ti[b]; });
    for (unsigned i = 0; i < cnt; ++i) out[i] = tmp[order[i]], out_inst[i] = ti[order[i]];
  }
  *n_out = cnt;
  return GVTB200_OK;
}

int gvtb200_framebuffer_clear(gvtb200_scene s) {
  NEED_INIT();
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  CU(cudaSetDevice(s->device));
  if (int rc = scene_fb(s)) return rc;
  CU(cudaMemsetAsync(s->d_fb, 0, s->fb_pix * 16, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return GVTB200_OK;
}
int gvtb200_framebuffer_clamp(gvtb200_scene s) {
  NEED_INIT();
  if (!s || !s->d_fb) return fail(GVTB200_ERR_INVALID, "no framebuffer");
  CU(cudaSetDevice(s->device));
  k_fb_clamp<<<blocks_for(s->fb_pix, 256), 256, 0, s->stream>>>(s->d_fb, s->fb_pix);
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  return GVTB200_OK;
}
int gvtb200_framebuffer_read(gvtb200_scene s, float *rgba, size_t n_floats) {
  NEED_INIT();
  if (!s || !s->d_fb || !rgba) return fail(GVTB200_ERR_INVALID, "no framebuffer / NULL argument");
  if (n_floats < s->fb_pix * 4) return fail(GVTB200_ERR_CAPACITY, "need %zu floats", s->fb_pix * 4);
  CU(cudaSetDevice(s->device));
  CU(cudaMemcpyAsync(rgba, s->d_fb, s->fb_pix * 16, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return GVTB200_OK;
}

// Parallel read-back: this rank's row tiles of the final image -- its own framebuffer (ImageScheduler: the rows it rendered)
// or rank `src_rank`'s (DomainScheduler: the root's composited image, read over NVLink) -- go to the same rows of a full-size
// host image over this rank's own PCIe link.  With one call per rank the host image (e.g. one shared pinned mapping) is
// complete without funnelling the film through one GPU.
int gvtb200_framebuffer_read_rows(gvtb200_scene s, float *rgba_image, int tile_rows, int tile_index, int tile_stride, int src_rank) {
  NEED_SCENE();
  if (!rgba_image || !s->d_fb) return fail(GVTB200_ERR_INVALID, "NULL image / no framebuffer");
  if (tile_rows < 1 || tile_stride < 1 || tile_index < 0 || tile_index >= tile_stride)
    return fail(GVTB200_ERR_INVALID, "bad tiling rows=%d index=%d stride=%d", tile_rows, tile_index, tile_stride);
  const float4 *src = (const float4 *)s->d_fb;
  if (src_rank >= 0) {
    if (!s->p2p.attached || src_rank >= s->p2p.world) return fail(GVTB200_ERR_INVALID, "rank %d is not an attached peer", src_rank);
    src = (const float4 *)s->p2p.peer_fb[src_rank];
  }
  RowTiles tiles;
  tiles.row0 = 0, tiles.row1 = s->H, tiles.tile_rows = tile_rows, tiles.index = tile_index, tiles.stride = tile_stride;
  const size_t rows = tiles_local_rows(tiles), n = rows * (size_t)s->W;
  if (!n) return GVTB200_OK;
  if (s->rows_cap < n) {
    dfree(s->d_rows);
    CU(dalloc(&s->d_rows, n));
    s->rows_cap = n;
  }
  cudaStream_t st = s->stream;
  k_rows_pack<<<std::min<unsigned>(g_dev[s->device].sms * 8, blocks_for(n, 256)), 256, 0, st>>>(src, s->d_rows, tiles, s->W);
  LAUNCHED(1);
  const size_t row_bytes = (size_t)s->W * 16, tile_bytes = row_bytes * tile_rows;
  const size_t full = rows / tile_rows, rem = rows - full * tile_rows;
  char *dst0 = (char *)rgba_image + (size_t)tile_index * tile_bytes;
  if (full) CU(cudaMemcpy2DAsync(dst0, tile_bytes * tile_stride, s->d_rows, tile_bytes, tile_bytes, full, cudaMemcpyDeviceToHost, st));
  if (rem) CU(cudaMemcpyAsync(dst0 + full * tile_bytes * tile_stride, (char *)s->d_rows + full * tile_bytes, rem * row_bytes, cudaMemcpyDeviceToHost, st));
  CU(cudaGetLastError());
  return GVTB200_OK;
}

// per-launch instrumentation: bracket one kernel launch with events, snapshot the counters after it
static int launch_rec_begin(gvtb200_scene s, int kind) {
  if (!s->launch_prof || s->n_launch_rec >= gvtb200_scene_s::MAX_LAUNCH_REC) return -1;
  int k = s->n_launch_rec++;
  s->launch_kind[k] = kind;
  for (int j = 0; j < 2; ++j)
    if (!s->lev[2 * k + j] && cudaEventCreate(&s->lev[2 * k + j]) != cudaSuccess) return -1;
  cudaEventRecord(s->lev[2 * k], s->stream);
  return k;
}
static void launch_rec_end(gvtb200_scene s, int k) {
  if (k < 0) return;
  cudaEventRecord(s->lev[2 * k + 1], s->stream);
  if (s->d_snap) cudaMemcpyAsync(s->d_snap + k, s->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToDevice, s->stream);
}

// one wave = Adapter::trace on every queued ray + the shadow rays it generated; `depth` waves drain a frame
static int run_waves(gvtb200_scene s, const FrameParams &fp, int waves, bool timed, const CameraGen *gen = nullptr,
                     const RayQueue *first_in = nullptr, const RayQueue *peer_inbox = nullptr) {
  SceneDev sd = scene_dev(*s);
  TraceQueues tq;
  ShadeQueues sq;
  memset(&tq, 0, sizeof(tq));
  tq.peer_inbox = peer_inbox;
  for (int wave = 0; wave < waves; ++wave) {
    tq.gen.enabled = 0;
    if (gen && wave == 0) tq.gen = *gen; // the first wave makes its own camera rays (no queue round trip through HBM)
    CU(cudaMemsetAsync(s->d_work, 0, 32, s->stream));
    CU(cudaMemsetAsync(s->q_b.q.count, 0, 4, s->stream));
    CU(cudaMemsetAsync(s->q_sh.q.count, 0, 4, s->stream));
    CU(cudaMemsetAsync(s->q_hit.q.count, 0, 4, s->stream));
    tq.in = (first_in && wave == 0) ? *first_in : s->q_a.q;
    tq.hq = s->q_hit.q, tq.hit = s->q_hit.hit, tq.mvq = null_queue(s->d_work + 2), tq.exq = s->q_ex.q;
    tq.fin = null_queue(s->d_work + 5), tq.work = s->d_work;
    if (timed && wave == waves - 1) CU(cudaEventRecord(s->ev[2], s->stream));
    int rk = launch_rec_begin(s, 1);
    launch_trace(s->device, s->profile, fp, sd, tq, 0, s->stream);
    launch_rec_end(s, rk);
    sq.hq = s->q_hit.q, sq.hit = s->q_hit.hit, sq.shq = s->q_sh.q, sq.nxq = s->q_b.q, sq.fin = null_queue(s->d_work + 5);
    rk = launch_rec_begin(s, 2);
    launch_shade(s->device, fp, sd, sq, 0, s->stream);
    launch_rec_end(s, rk);
    if (timed && wave == waves - 1) CU(cudaEventRecord(s->ev[3], s->stream));
    tq.gen.enabled = 0;
    tq.in = s->q_sh.q, tq.hq = null_queue(s->d_work + 3), tq.hit = nullptr, tq.work = s->d_work + 1;
    rk = launch_rec_begin(s, 3);
    launch_trace(s->device, s->profile, fp, sd, tq, 0, s->stream);
    launch_rec_end(s, rk);
    if (timed && wave == waves - 1) CU(cudaEventRecord(s->ev[4], s->stream));
    std::swap(s->q_a, s->q_b);
  }
  return 0;
}

static int frame_queues(gvtb200_scene s, size_t count) {
  size_t nl = std::max<size_t>(1, s->lights.size());
  if (count * nl > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "frame has too many rays for one queue");
  auto lim = [&](size_t n) { return s->queue_limit ? std::min(n, s->queue_limit) : n; };
  if (s->queue_limit) // a smaller budget than what is allocated: start over
    for (QueueMem *q : { &s->q_a, &s->q_b, &s->q_sh, &s->q_hit, &s->q_ex })
      if (q->slab && q->q.cap > std::max<size_t>(32, s->queue_limit * nl)) q->release();
  CU(s->q_a.ensure(lim(count)));
  CU(s->q_b.ensure(s->depth > 1 ? lim(count) : 32));
  CU(s->q_sh.ensure(lim(count * nl)));
  CU(s->q_hit.ensure(lim(count), true));
  CU(s->q_ex.ensure(s->world > 1 ? lim(count) : 32));
  return 0;
}

// p2p_root < 0: plain frame.  p2p_root >= 0: ImageScheduler over NVLink -- the root releases its peers after its own clear,
// every rank clamps and stores its rows into the root's framebuffer, the root waits for all of them.
// phases (peer path only): 1 = clear (+ the root's go), 2 = trace + store the rows, 4 = the root's wait.  A production call
// runs all three; the single-process tests, where every simulated rank shares one GPU, run the root's phase 1 before the
// peers' frames and its phases 2|4 after them so that no spinning kernel ever waits for a kernel that is not launched yet.
static int frame_tiles(gvtb200_scene s, int row0, int row1, int tile_rows, int tile_index, int tile_stride, int p2p_root, int phases = 7) {
  NEED_SCENE();
  if (s->world > 1) return fail(GVTB200_ERR_INVALID, "domain-partitioned scene: use the gvtb200_domain_* calls");
  if (row0 < 0 || row1 > s->H || row0 > row1) return fail(GVTB200_ERR_INVALID, "bad row range [%d,%d) of %d", row0, row1, s->H);
  if (tile_rows < 1 || tile_stride < 1 || tile_index < 0 || tile_index >= tile_stride)
    return fail(GVTB200_ERR_INVALID, "bad tiling rows=%d index=%d stride=%d", tile_rows, tile_index, tile_stride);
  if (int rc = scene_fb(s)) return rc;
  RowTiles tiles;
  tiles.row0 = row0, tiles.row1 = row1, tiles.tile_rows = tile_rows, tiles.index = tile_index, tiles.stride = tile_stride;
  const size_t s2 = (size_t)s->samples * s->samples;
  const size_t count = tiles_local_rows(tiles) * s->W * s2;
  if (int rc = frame_queues(s, count)) return rc;
  FrameParams fp;
  build_frame_params(*s, fp);
  cudaStream_t st = s->stream;
  const bool p2p = p2p_root >= 0;
  if (!p2p) phases = 7;
  if (phases & 1) {
    CU(cudaEventRecord(s->ev[0], st));
    CU(cudaMemsetAsync(s->d_fb, 0, s->fb_pix * 16, st)); // clearBuffer, TracerBase.h:310
    CU(cudaMemsetAsync(s->d_ctr, 0, sizeof(DevCounters), st));
    CU(cudaMemsetAsync(s->q_a.q.count, 0, 4, st));
    CU(cudaMemsetAsync(s->q_ex.q.count, 0, 4, st));
    s->n_launch_rec = 0;
    if (p2p) ++s->p2p.seq;
    if (p2p && s->p2p.rank == p2p_root) { // the framebuffer is clear: the peers may store their rows of frame `seq`
      k_p2p_go<<<1, GVT_P2P_MAX, 0, st>>>(s->p2p.mail, s->p2p.world, s->p2p.rank, s->p2p.seq);
      LAUNCHED(1);
    }
  }
  const unsigned seq = s->p2p.seq;
  if (count && (phases & 2)) {
    // few instances: the first trace wave makes and routes its own camera rays in registers.  Many instances: the
    // top-level routing is long and runs better with full warps in its own kernel (measured: 32x32 grid, 1.79 vs 1.87 ms)
    if (s->inst.size() <= 4) {
      CameraGen gen;
      gen.enabled = 1, gen.tiles = tiles, gen.count = (unsigned)count;
      if (int rc = run_waves(s, fp, std::max(1, s->depth), true, &gen)) return rc;
    } else {
      int rk = launch_rec_begin(s, 0);
      k_camera_route<<<g_dev[s->device].sms * 8, 256, 0, st>>>(fp, scene_dev(*s), tiles, count, s->q_a.q);
      LAUNCHED(1);
      launch_rec_end(s, rk);
      if (int rc = run_waves(s, fp, std::max(1, s->depth), true)) return rc;
    }
  }
  if (!p2p) {
    k_fb_clamp<<<blocks_for(s->fb_pix, 256), 256, 0, st>>>(s->d_fb, s->fb_pix);
    LAUNCHED(1);
  } else {
    const int rank = s->p2p.rank;
    Mailbox *mine = s->p2p.d_mail, *root_mail = s->p2p.mail.m[p2p_root];
    const unsigned grid = (unsigned)std::min<size_t>((size_t)g_dev[s->device].sms * 8, std::max<size_t>(1, (count / std::max<size_t>(1, s2) + 255) / 256));
    if (rank == p2p_root) {
      if (count && (phases & 2)) {
        k_rows_to_root<<<grid, 256, 0, st>>>((const float4 *)s->d_fb, (float4 *)s->d_fb, tiles, s->W, mine, nullptr, rank, seq);
        LAUNCHED(1);
      }
      if (phases & 4) {
        k_p2p_wait<<<1, GVT_P2P_MAX, 0, st>>>(mine->done, s->p2p.world, rank, seq, &mine->err, g_peer_timeout_ns);
        LAUNCHED(1);
      }
    } else if (phases & 2) {
      k_p2p_wait<<<1, 32, 0, st>>>(&mine->go, 1, -1, seq, &mine->err, g_peer_timeout_ns);
      k_rows_to_root<<<grid, 256, 0, st>>>((const float4 *)s->d_fb, (float4 *)s->p2p.peer_fb[p2p_root], tiles, s->W, mine, root_mail, rank, seq);
      LAUNCHED(2);
    }
  }
  CU(cudaEventRecord(s->ev[1], st));
  CU(cudaGetLastError());
  s->timing_valid = count > 0;
  return GVTB200_OK;
}
int gvtb200_render_frame_tiles(gvtb200_scene s, int row0, int row1, int tile_rows, int tile_index, int tile_stride) {
  return frame_tiles(s, row0, row1, tile_rows, tile_index, tile_stride, -1);
}
int gvtb200_render_frame_async(gvtb200_scene s, int row0, int row1) { return gvtb200_render_frame_tiles(s, row0, row1, 1, 0, 1); }
// run the scene's work on a caller-owned CUDA stream (e.g. the stream NCCL collectives are enqueued on); NULL restores
// the scene's own stream
int gvtb200_scene_set_stream(gvtb200_scene s, void *cuda_stream) {
  NEED_INIT();
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  CU(cudaStreamSynchronize(s->stream));
  s->stream = cuda_stream ? (cudaStream_t)cuda_stream : s->own_stream;
  return GVTB200_OK;
}
unsigned long long gvtb200_launch_count(void) { return g_launches; }
// The reference never drops a ray (TracerBase.h:403-409 grows its queues): a frame during which any queue append found its
// queue full, or a traversal stack overflowed, is reported as an error instead of a silently incomplete image.
static int check_frame_errors(gvtb200_scene s) {
  unsigned long long ov[2] = { 0, 0 }; // stack_overflows, queue_overflows are adjacent
  CU(cudaMemcpyAsync(ov, &s->d_ctr->stack_overflows, sizeof(ov), cudaMemcpyDeviceToHost, s->stream));
  unsigned perr = 0, merr = 0;
  if (s->p2p.d_mail) {
    CU(cudaMemcpyAsync(&perr, &s->p2p.d_mail->pc.err, 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(&merr, &s->p2p.d_mail->err, 4, cudaMemcpyDeviceToHost, s->stream));
  }
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  if (merr || (perr & 1u)) {
    if (merr) CU(cudaMemset(&s->p2p.d_mail->err, 0, 4));
    if (perr) CU(cudaMemset(&s->p2p.d_mail->pc.err, 0, 4));
    return fail(GVTB200_ERR_CUDA, "NVLink peer path: rank %d waited more than %.1f s for a peer (frame %u)", s->p2p.rank, (double)g_peer_timeout_ns * 1e-9, s->p2p.seq);
  }
  if (ov[1] || (perr & 2u)) {
    if (perr) CU(cudaMemset(&s->p2p.d_mail->pc.err, 0, 4));
    return fail(GVTB200_ERR_CAPACITY, "%llu rays found a ray queue full during the frame (the image is incomplete)", ov[1]);
  }
  if (ov[0]) return fail(GVTB200_ERR_CAPACITY, "the traversal stack overflowed (GVT_BLAS_STACK = %d entries is too small for this BVH)", GVT_BLAS_STACK);
  return GVTB200_OK;
}
int gvtb200_scene_sync(gvtb200_scene s) {
  NEED_INIT();
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  CU(cudaSetDevice(s->device));
  return check_frame_errors(s);
}
int gvtb200_scene_counters(gvtb200_scene s, gvtb200_counters *out) {
  static_assert(sizeof(gvtb200_counters) == sizeof(DevCounters), "counter layouts must match");
  NEED_INIT();
  if (!s || !out) return fail(GVTB200_ERR_INVALID, "NULL argument");
  CU(cudaMemcpyAsync(out, s->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return GVTB200_OK;
}
int gvtb200_render_frame(gvtb200_scene s, int row0, int row1, float *rgba_out, gvtb200_counters *counters) {
  int rc = gvtb200_render_frame_async(s, row0, row1);
  if (rc) return rc;
  if (rgba_out) CU(cudaMemcpyAsync(rgba_out, s->d_fb, s->fb_pix * 16, cudaMemcpyDeviceToHost, s->stream));
  if (counters) CU(cudaMemcpyAsync(counters, s->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, s->stream));
  return check_frame_errors(s);
}
int gvtb200_scene_profile_launches(gvtb200_scene s, int enable) {
  NEED_INIT();
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  CU(cudaSetDevice(s->device));
  if (enable && !s->d_snap) CU(dalloc(&s->d_snap, (size_t)gvtb200_scene_s::MAX_LAUNCH_REC));
  s->launch_prof = enable != 0;
  s->n_launch_rec = 0;
  return GVTB200_OK;
}
int gvtb200_scene_launch_records(gvtb200_scene s, int cap, int *n, int32_t *kind, float *ms, gvtb200_counters *after) {
  NEED_INIT();
  if (!s || !n) return fail(GVTB200_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(s->device));
  CU(cudaStreamSynchronize(s->stream));
  int k = std::min(cap, s->n_launch_rec);
  for (int i = 0; i < k; ++i) {
    if (kind) kind[i] = s->launch_kind[i];
    if (ms) CU(cudaEventElapsedTime(&ms[i], s->lev[2 * i], s->lev[2 * i + 1]));
  }
  if (after && k) CU(cudaMemcpy(after, s->d_snap, (size_t)k * sizeof(DevCounters), cudaMemcpyDeviceToHost));
  *n = k;
  return GVTB200_OK;
}
int gvtb200_scene_profile_counters(gvtb200_scene s, int enable) {
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  s->profile = enable != 0;
  return GVTB200_OK;
}
int gvtb200_scene_device_framebuffer(gvtb200_scene s, void **d_rgba, void **cuda_stream) {
  NEED_INIT();
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  CU(cudaSetDevice(s->device));
  if (int rc = scene_fb(s)) return rc;
  if (d_rgba) *d_rgba = s->d_fb;
  if (cuda_stream) *cuda_stream = (void *)s->stream;
  return GVTB200_OK;
}
int gvtb200_scene_last_timing(gvtb200_scene s, float *frame_ms, float kernel_ms[3]) {
  NEED_INIT();
  if (!s || !s->timing_valid) return fail(GVTB200_ERR_INVALID, "no timed frame");
  CU(cudaEventSynchronize(s->ev[1]));
  float a = 0, b = 0, c = 0, d = 0;
  CU(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
  CU(cudaEventElapsedTime(&b, s->ev[0], s->ev[2]));
  CU(cudaEventElapsedTime(&c, s->ev[2], s->ev[3]));
  CU(cudaEventElapsedTime(&d, s->ev[3], s->ev[4]));
  if (frame_ms) *frame_ms = a;
  if (kernel_ms) kernel_ms[0] = b, kernel_ms[1] = c, kernel_ms[2] = d;
  return GVTB200_OK;
}

// =============================================================================================================
// Tracer<DomainScheduler>
// =============================================================================================================
int gvtb200_domain_set_owner(gvtb200_scene s, const int32_t *owner, int n, int rank, int world) {
  if (!s || !owner) return fail(GVTB200_ERR_INVALID, "NULL argument");
  if (n != (int)s->inst.size()) return fail(GVTB200_ERR_INVALID, "owner table has %d entries, scene has %zu instances", n, s->inst.size());
  if (world < 1 || world > 1024 || rank < 0 || rank >= world) return fail(GVTB200_ERR_INVALID, "bad rank/world %d/%d", rank, world);
  for (int i = 0; i < n; ++i) {
    if (owner[i] < 0 || owner[i] >= world) return fail(GVTB200_ERR_INVALID, "owner[%d] = %d outside 0..%d", i, owner[i], world - 1);
    s->inst[i].owner = owner[i];
  }
  s->rank = rank, s->world = world;
  s->committed = false;
  return GVTB200_OK;
}
int gvtb200_domain_begin_frame(gvtb200_scene s) {
  NEED_SCENE();
  if (int rc = scene_fb(s)) return rc;
  const size_t count = (size_t)s->W * s->H * s->samples * s->samples;
  if (int rc = frame_queues(s, count)) return rc;
  FrameParams fp;
  build_frame_params(*s, fp);
  cudaStream_t st = s->stream;
  CU(cudaMemsetAsync(s->d_fb, 0, s->fb_pix * 16, st));
  CU(cudaMemsetAsync(s->d_ctr, 0, sizeof(DevCounters), st));
  CU(cudaMemsetAsync(s->q_a.q.count, 0, 4, st));
  CU(cudaMemsetAsync(s->q_ex.q.count, 0, 4, st));
  // every rank generates all camera rays and keeps those whose first instance it owns (DomainTracer.h:148-181)
  RowTiles tiles;
  tiles.row0 = 0, tiles.row1 = s->H, tiles.tile_rows = 1, tiles.index = 0, tiles.stride = 1;
  k_camera_route<<<g_dev[s->device].sms * 8, 256, 0, st>>>(fp, scene_dev(*s), tiles, count, s->q_a.q);
  LAUNCHED(1);
  CU(cudaGetLastError());
  s->timing_valid = false;
  return GVTB200_OK;
}
int gvtb200_domain_drain_local(gvtb200_scene s) {
  NEED_SCENE();
  FrameParams fp;
  build_frame_params(*s, fp);
  if (int rc = run_waves(s, fp, std::max(1, s->depth), false)) return rc;
  CU(cudaGetLastError());
  return GVTB200_OK;
}
int gvtb200_domain_export(gvtb200_scene s, uint32_t *counts, void **d_export, size_t *n_total) {
  NEED_SCENE();
  if (!counts || !d_export || !n_total) return fail(GVTB200_ERR_INVALID, "NULL argument");
  cudaStream_t st = s->stream;
  const int P = s->world;
  unsigned n_ex = 0;
  CU(cudaMemcpyAsync(&n_ex, s->q_ex.q.count, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemsetAsync(s->d_peer, 0, 4 * (size_t)P, st));
  CU(cudaStreamSynchronize(st));
  if (n_ex > s->q_ex.q.cap) return fail(GVTB200_ERR_CAPACITY, "export queue overflow (%u > %u)", n_ex, s->q_ex.q.cap);
  std::vector<unsigned> c(P, 0), off(P, 0);
  if (n_ex) {
    k_export_count<<<g_dev[s->device].sms * 4, 256, 0, st>>>(s->q_ex.q, s->d_inst, s->d_peer);
    CU(cudaMemcpyAsync(c.data(), s->d_peer, 4 * (size_t)P, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int p = 1; p < P; ++p) off[p] = off[p - 1] + c[p - 1];
    if (s->export_cap < n_ex) {
      dfree(s->d_export);
      CU(dalloc(&s->d_export, 5 * (size_t)n_ex));
      s->export_cap = n_ex;
    }
    CU(cudaMemcpyAsync(s->d_peer, off.data(), 4 * (size_t)P, cudaMemcpyHostToDevice, st));
    k_export_pack<<<g_dev[s->device].sms * 4, 256, 0, st>>>(s->q_ex.q, s->d_inst, s->d_peer, s->d_export);
    CU(cudaMemsetAsync(s->q_ex.q.count, 0, 4, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
  }
  for (int p = 0; p < P; ++p) counts[p] = c[p];
  *d_export = s->d_export;
  *n_total = n_ex;
  return GVTB200_OK;
}
int gvtb200_domain_import(gvtb200_scene s, const void *d_rays, size_t n_rays) {
  NEED_SCENE();
  if (n_rays == 0) return GVTB200_OK;
  if (!d_rays) return fail(GVTB200_ERR_INVALID, "NULL argument");
  cudaStream_t st = s->stream;
  unsigned have = 0;
  CU(cudaMemcpyAsync(&have, s->q_a.q.count, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  have = std::min(have, s->q_a.q.cap);
  if ((size_t)have + n_rays > s->q_a.q.cap) {
    if (have) return fail(GVTB200_ERR_CAPACITY, "local queue full (%u + %zu > %u)", have, n_rays, s->q_a.q.cap);
    CU(s->q_a.ensure(n_rays));
    CU(s->q_b.ensure(s->depth > 1 ? n_rays : 32));
    CU(s->q_sh.ensure(n_rays * std::max<size_t>(1, s->lights.size())));
    CU(s->q_hit.ensure(n_rays, true));
  }
  k_import<<<blocks_for(n_rays, 256), 256, 0, st>>>((const uint4 *)d_rays, (unsigned)n_rays, s->q_a.q, have);
  unsigned total = have + (unsigned)n_rays;
  CU(cudaMemcpyAsync(s->q_a.q.count, &total, 4, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  return GVTB200_OK;
}
int gvtb200_domain_pending(gvtb200_scene s, uint64_t *n) {
  NEED_SCENE();
  if (!n) return fail(GVTB200_ERR_INVALID, "NULL argument");
  unsigned a = 0, e = 0;
  CU(cudaMemcpyAsync(&a, s->q_a.q.count, 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(&e, s->q_ex.q.count, 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  *n = (uint64_t)a + e;
  return GVTB200_OK;
}

// =============================================================================================================
// NVLink peer path
// =============================================================================================================
static RayQueue queue_at(void *slab, size_t cap) { // the layout QueueMem::ensure gives an inbound slab (no hit column, ready column)
  RayQueue q;
  char *p = (char *)slab;
  q.count = (unsigned *)p, p += 256;
  q.c0 = (float4 *)p, p += cap * 16;
  q.c1 = (float4 *)p, p += cap * 16;
  q.c2 = (float4 *)p, p += cap * 16;
  q.c3 = (int4 *)p, p += cap * 16;
  q.inst = (int *)p, p += cap * 4;
  q.ready = (unsigned *)p;
  q.cap = (unsigned)cap;
  return q;
}
// capacity of an inbound queue: what one exchange round can bring (every camera ray with its shadow rays and its bounce),
// and, for the persistent scheduler, everything a rank is handed in a whole frame
static size_t inbox_capacity(const gvtb200_scene_s *s) {
  const size_t nl = std::max<size_t>(1, s->lights.size());
  const size_t frame = (size_t)s->W * s->H * s->samples * s->samples;
  const size_t cap = std::min<size_t>(frame * (nl + 1) * (size_t)std::max(1, std::min(s->depth, 4)), (size_t)1 << 27);
  return s->queue_limit ? std::max<size_t>(64, std::min(cap, s->queue_limit)) : cap;
}
static int p2p_own_objects(gvtb200_scene s, bool both_inboxes) {
  if (int rc = scene_fb(s)) return rc;
  const size_t cap = inbox_capacity(s);
  if (cap > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "frame has too many rays for one queue");
  for (int k = 0; k < 2; ++k) {
    if (s->p2p.inbox[k].slab && s->p2p.inbox[k].q.cap != cap) s->p2p.inbox[k].release();
    const bool used = s->world > 1 ? (k == 0 || both_inboxes) : (k == 0 && !both_inboxes); // only the DomainScheduler exchanges rays
    CU(s->p2p.inbox[k].ensure(used ? cap : 32, false, true));
  }
  if (!s->p2p.d_mail) CU(dalloc(&s->p2p.d_mail, 1));
  CU(cudaMemset(s->p2p.d_mail, 0, sizeof(Mailbox)));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
int gvtb200_p2p_export(gvtb200_scene s, int world, gvtb200_peer_desc *out) {
  NEED_SCENE();
  if (!out) return fail(GVTB200_ERR_INVALID, "out is NULL");
  if (world < 1 || world > GVT_P2P_MAX) return fail(GVTB200_ERR_INVALID, "world size %d outside 1..%d", world, GVT_P2P_MAX);
  if (s->p2p.attached) return fail(GVTB200_ERR_INVALID, "already attached: gvtb200_p2p_detach first");
  if (s->p2p.private_box) gvtb200_p2p_detach(s);
  if (int rc = p2p_own_objects(s, true)) return rc;
  memset(out, 0, sizeof(*out));
  void *ptrs[4] = { s->d_fb, s->p2p.inbox[0].slab, s->p2p.inbox[1].slab, s->p2p.d_mail };
  uint8_t *handles[4] = { out->fb_handle, out->inbox_handle[0], out->inbox_handle[1], out->mail_handle };
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  for (int k = 0; k < 4; ++k) {
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ptrs[k]));
    memcpy(handles[k], &h, 64);
    out->raw[k] = (uint64_t)(uintptr_t)ptrs[k];
  }
  out->inbox_cap = s->p2p.inbox[0].q.cap, out->inbox_cap1 = s->p2p.inbox[1].q.cap, out->fb_pix = s->fb_pix, out->pid = (int32_t)getpid(), out->device = s->device;
  s->p2p.exported = true;
  return GVTB200_OK;
}
// another device of this process: direct peer access (one host thread per GPU, no IPC)
static int enable_peer(int self, int other) {
  if (self == other) return 0;
  int can = 0;
  CU(cudaDeviceCanAccessPeer(&can, self, other));
  if (!can) return fail(GVTB200_ERR_INVALID, "device %d cannot access device %d (no NVLink / PCIe peer path)", self, other);
  cudaError_t e = cudaDeviceEnablePeerAccess(other, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(GVTB200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", other, cudaGetErrorString(e));
  (void)cudaGetLastError();
  return 0;
}
int gvtb200_p2p_attach(gvtb200_scene s, int rank, int world, const gvtb200_peer_desc *all) {
  NEED_SCENE();
  if (!all) return fail(GVTB200_ERR_INVALID, "NULL argument");
  if (!s->p2p.exported) return fail(GVTB200_ERR_INVALID, "gvtb200_p2p_export has not been called on this scene");
  if (world < 1 || world > GVT_P2P_MAX || rank < 0 || rank >= world) return fail(GVTB200_ERR_INVALID, "bad rank/world %d/%d", rank, world);
  if (s->world > 1 && (s->world != world || s->rank != rank)) return fail(GVTB200_ERR_INVALID, "rank/world differ from gvtb200_domain_set_owner (%d/%d)", s->rank, s->world);
  if (s->p2p.attached) return fail(GVTB200_ERR_INVALID, "already attached");
  std::vector<RayQueue> table(2 * (size_t)world);
  for (int r = 0; r < world; ++r) {
    const gvtb200_peer_desc &d = all[r];
    if (d.fb_pix != s->fb_pix) return fail(GVTB200_ERR_INVALID, "rank %d renders a different film (%llu vs %zu pixels)", r, (unsigned long long)d.fb_pix, s->fb_pix);
    void *ptrs[4];
    if (r == rank) {
      ptrs[0] = s->d_fb, ptrs[1] = s->p2p.inbox[0].slab, ptrs[2] = s->p2p.inbox[1].slab, ptrs[3] = s->p2p.d_mail;
    } else if (d.pid == (int32_t)getpid()) {
      // another scene of this process: same address space.  The same device (several simulated ranks on one GPU: the
      // single-process tests) needs nothing; another device (one host thread per GPU) needs peer access.
      if (int rc = enable_peer(s->device, d.device)) return rc;
      for (int k = 0; k < 4; ++k) ptrs[k] = (void *)(uintptr_t)d.raw[k];
    } else {
      const uint8_t *handles[4] = { d.fb_handle, d.inbox_handle[0], d.inbox_handle[1], d.mail_handle };
      for (int k = 0; k < 4; ++k) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles[k], 64);
        cudaError_t e = cudaIpcOpenMemHandle(&ptrs[k], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          gvtb200_p2p_detach(s);
          return fail(GVTB200_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d, object %d): %s", r, k, cudaGetErrorString(e));
        }
        s->p2p.opened[r][k] = ptrs[k];
      }
    }
    s->p2p.peer_fb[r] = (float *)ptrs[0];
    table[r] = queue_at(ptrs[1], d.inbox_cap), table[(size_t)world + r] = queue_at(ptrs[2], d.inbox_cap1);
    s->p2p.mail.m[r] = (Mailbox *)ptrs[3];
  }
  dfree(s->p2p.d_peer_inbox);
  CU(dalloc(&s->p2p.d_peer_inbox, table.size()));
  CU(cudaMemcpy(s->p2p.d_peer_inbox, table.data(), table.size() * sizeof(RayQueue), cudaMemcpyHostToDevice));
  s->p2p.rank = rank, s->p2p.world = world, s->p2p.seq = 0, s->p2p.attached = true;
  return GVTB200_OK;
}
int gvtb200_p2p_detach(gvtb200_scene s) {
  if (!s) return GVTB200_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (auto &r : s->p2p.opened)
    for (auto &p : r) {
      if (p) cudaIpcCloseMemHandle(p);
      p = nullptr;
    }
  dfree(s->p2p.d_peer_inbox), dfree(s->p2p.d_mail), dfree(s->p2p.d_inst_load);
  s->p2p.hitq.release();
  s->p2p.inbox[0].release(), s->p2p.inbox[1].release();
  s->p2p.exported = s->p2p.attached = s->p2p.private_box = false;
  return GVTB200_OK;
}
int gvtb200_image_frame_p2p(gvtb200_scene s, int row0, int row1, int tile_rows, int root, int phases) {
  if (!s || !s->p2p.attached) return fail(GVTB200_ERR_INVALID, "scene is not attached to its peers (gvtb200_p2p_attach)");
  if (root < 0 || root >= s->p2p.world) return fail(GVTB200_ERR_INVALID, "root %d outside 0..%d", root, s->p2p.world - 1);
  if (phases < 1 || phases > 7) return fail(GVTB200_ERR_INVALID, "phases %d outside 1..7", phases);
  return frame_tiles(s, row0, row1, tile_rows, s->p2p.rank, s->p2p.world, root, phases);
}
int gvtb200_domain_round_p2p(gvtb200_scene s, int round, void **d_exported_total) {
  NEED_SCENE();
  if (!s->p2p.attached || s->world < 2) return fail(GVTB200_ERR_INVALID, "needs gvtb200_domain_set_owner with world > 1 and gvtb200_p2p_attach");
  if (round < 0) return fail(GVTB200_ERR_INVALID, "round %d", round);
  // the frame's queues must hold whatever the peers may send: up to the inbound queue's capacity
  const size_t cap = s->p2p.inbox[0].q.cap, nl = std::max<size_t>(1, s->lights.size());
  if (cap * nl > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "frame has too many rays for one queue");
  CU(s->q_b.ensure(s->depth > 1 ? cap : 32));
  CU(s->q_sh.ensure(cap * nl));
  CU(s->q_hit.ensure(cap, true));
  FrameParams fp;
  build_frame_params(*s, fp);
  const RayQueue *first = round == 0 ? nullptr : &s->p2p.inbox[(round - 1) & 1].q;
  const RayQueue *peers = s->p2p.d_peer_inbox + (size_t)(round & 1) * s->p2p.world;
  if (int rc = run_waves(s, fp, std::max(1, s->depth), false, nullptr, first, peers)) return rc;
  // what this round consumed is empty again before any peer can store into it (round + 1 starts after the all-reduce)
  if (round > 0) CU(cudaMemsetAsync(s->p2p.inbox[(round - 1) & 1].q.count, 0, 4, s->stream));
  CU(cudaGetLastError());
  if (d_exported_total) *d_exported_total = (void *)&s->d_ctr->rays_exported;
  return GVTB200_OK;
}
// ---- persistent DomainScheduler -----------------------------------------------------------------------------------------
int gvtb200_domain_set_servers(gvtb200_scene s, const int32_t *ranks, const float *weights, int n) {
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  if (!ranks) { // back to "the owner serves its instances"
    s->p2p.servers.clear();
    cudaSetDevice(s->device);
    dfree(s->p2p.d_servers);
    return GVTB200_OK;
  }
  if (n != (int)s->inst.size()) return fail(GVTB200_ERR_INVALID, "server table has %d entries, scene has %zu instances", n, s->inst.size());
  std::vector<uint4> tab((size_t)n);
  for (int i = 0; i < n; ++i) {
    unsigned packed = 0, cnt = 0;
    double w[4] = { 0, 0, 0, 0 }, sum = 0;
    for (int j = 0; j < 4; ++j) {
      const int r = ranks[4 * i + j];
      if (r < 0) continue;
      if (r >= std::max(1, s->world) || r > 255) return fail(GVTB200_ERR_INVALID, "instance %d: server rank %d outside 0..%d", i, r, s->world - 1);
      const double wj = weights ? (double)weights[4 * i + j] : 1.0;
      if (!(wj > 0.0)) continue;
      if (r == s->rank && !s->inst[i].mesh) return fail(GVTB200_ERR_INVALID, "instance %d: this rank is listed as a server but holds no BLAS for it", i);
      packed |= (unsigned)r << (8 * cnt), w[cnt] = wj, sum += wj, ++cnt;
    }
    if (!cnt) return fail(GVTB200_ERR_INVALID, "instance %d has no server", i);
    unsigned thr[3] = { 65536u, 65536u, 65536u };
    double acc = 0;
    for (unsigned j = 0; j + 1 < cnt; ++j) acc += w[j] / sum, thr[j] = (unsigned)std::min(65535.0, std::max(0.0, acc * 65536.0 + 0.5));
    // the 16-bit hash h picks server #(thresholds <= h): unused thresholds are above every h
    tab[i] = make_uint4(packed, std::min(thr[0], 65535u) | (std::min(thr[1], 65535u) << 16), thr[2], cnt);
    if (cnt < 2) tab[i].y = 0xFFFFFFFFu;
    else if (cnt < 3) tab[i].y |= 0xFFFF0000u;
  }
  s->p2p.servers = tab;
  CU(cudaSetDevice(s->device));
  dfree(s->p2p.d_servers);
  CU(dalloc(&s->p2p.d_servers, tab.size()));
  CU(cudaMemcpy(s->p2p.d_servers, tab.data(), tab.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  return GVTB200_OK;
}
int gvtb200_scene_set_instance_mesh(gvtb200_scene s, int instance, gvtb200_mesh mesh) {
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  if (instance < 0 || instance >= (int)s->inst.size()) return fail(GVTB200_ERR_INVALID, "instance %d outside 0..%zu", instance, s->inst.size());
  if (mesh && mesh->device != s->device) return fail(GVTB200_ERR_INVALID, "the mesh lives on device %d, the scene on device %d", mesh->device, s->device);
  s->inst[instance].mesh = mesh;
  s->committed = false;
  return GVTB200_OK;
}
static int persist_prepare(gvtb200_scene s);
int gvtb200_domain_persist_config(gvtb200_scene s, int grid_ctas, int shade_sms) {
  if (!s) return fail(GVTB200_ERR_INVALID, "scene is NULL");
  if (grid_ctas < 0 || shade_sms < 0) return fail(GVTB200_ERR_INVALID, "grid_ctas %d, shade_sms %d must be >= 0", grid_ctas, shade_sms);
  s->p2p.grid_limit = grid_ctas, s->p2p.shade_sms = shade_sms;
  NEED_SCENE();
  return persist_prepare(s);
}
static int persist_prepare(gvtb200_scene s) {
  const size_t cap = inbox_capacity(s);
  if (s->p2p.hitq.slab && s->p2p.hitq.q.cap != cap) s->p2p.hitq.release();
  CU(s->p2p.hitq.ensure(cap, true, true));
  if (!s->p2p.d_inst_load) CU(dalloc(&s->p2p.d_inst_load, std::max<size_t>(1, s->inst.size())));
  return 0;
}
int gvtb200_domain_frame_persist(gvtb200_scene s, int root, int tile_rows) {
  NEED_SCENE();
  const int world = std::max(1, s->world);
  if (world > 1 && !s->p2p.attached) return fail(GVTB200_ERR_INVALID, "needs gvtb200_p2p_attach (world > 1)");
  if (world == 1 && !s->p2p.attached) { // a private mailbox + inbox: one rank serves everything
    if (!s->p2p.private_box || s->p2p.inbox[0].q.cap != inbox_capacity(s)) {
      gvtb200_p2p_detach(s);
      if (int rc = p2p_own_objects(s, false)) return rc;
      RayQueue self = s->p2p.inbox[0].q;
      CU(dalloc(&s->p2p.d_peer_inbox, 1));
      CU(cudaMemcpy(s->p2p.d_peer_inbox, &self, sizeof(RayQueue), cudaMemcpyHostToDevice));
      memset(&s->p2p.mail, 0, sizeof(s->p2p.mail));
      s->p2p.mail.m[0] = s->p2p.d_mail, s->p2p.peer_fb[0] = s->d_fb;
      s->p2p.rank = 0, s->p2p.world = 1, s->p2p.seq = 0, s->p2p.private_box = true;
    }
    root = 0;
  }
  if (s->p2p.world != world) return fail(GVTB200_ERR_INVALID, "attached to %d ranks, the scene is partitioned over %d", s->p2p.world, world);
  if (root < 0 || root >= world) return fail(GVTB200_ERR_INVALID, "root %d outside 0..%d", root, world - 1);
  if (tile_rows < 1) return fail(GVTB200_ERR_INVALID, "tile_rows %d", tile_rows);
  if (!s->p2p.inbox[0].q.ready || s->p2p.inbox[0].q.cap < 64) return fail(GVTB200_ERR_INVALID, "no inbound queue (ImageScheduler attachment?)");
  const int me = s->p2p.rank;
  cudaStream_t st = s->stream;
  const bool count = s->profile;
  // (allocations happen here only the first time: cudaMalloc / cudaFree may wait for running kernels, and another rank's
  // persistent kernel on the same device -- the single-process tests -- would be waiting for this rank's launch)
  if (int rc = persist_prepare(s)) return rc;
  int grid = persist_grid(s->device, count);
  if (s->p2p.grid_limit > 0) grid = std::min(grid, s->p2p.grid_limit);
  grid = std::max(grid, 2);
  if (count) CU(cudaMemsetAsync(s->p2p.d_inst_load, 0, std::max<size_t>(1, s->inst.size()) * 8, st));
  if (!s->p2p.servers.empty() && s->p2p.servers.size() != s->inst.size()) return fail(GVTB200_ERR_INVALID, "server table is stale (instances changed)");
  FrameParams fp;
  build_frame_params(*s, fp);
  PersistParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.inbox = s->p2p.inbox[0].q;
  pp.peer_inbox = s->p2p.d_peer_inbox; // entries [0, world): the round-0 inbound queues
  pp.mail = s->p2p.mail;
  pp.root_fb = s->p2p.peer_fb[root];
  pp.hitq = s->p2p.hitq.q, pp.hit = s->p2p.hitq.hit;
  pp.servers = s->p2p.d_servers;
  pp.inst_load = s->p2p.d_inst_load;
  pp.timeout_ns = g_peer_timeout_ns;
  pp.tiles.row0 = 0, pp.tiles.row1 = s->H, pp.tiles.tile_rows = tile_rows, pp.tiles.index = me, pp.tiles.stride = world;
  const size_t cam = tiles_local_rows(pp.tiles) * (size_t)s->W * s->samples * s->samples;
  if (cam > 0x7FFFFFFFu) return fail(GVTB200_ERR_INVALID, "frame has too many camera rays");
  pp.cam_count = (unsigned)cam;
  pp.seq = ++s->p2p.seq;
  pp.rank = me, pp.world = world, pp.root = root;
  // roles: with the GPU to itself the kernel splits by SM (one role's code per instruction cache); a bounded grid (several
  // simulated ranks on one GPU) cannot count on where its CTAs land and splits by CTA index instead
  const int full = persist_grid(s->device, count), sms = g_dev[s->device].sms;
  const int want = s->p2p.shade_sms > 0 ? s->p2p.shade_sms : std::max(2, sms / 10);
  pp.shade_sms = grid == full ? std::min(want, sms - 1) : 0;
  pp.shade_mod = std::max(2, std::min(grid, sms / std::max(1, want)));
  pp.clamp = 1, pp.timeline = 1;
  pp.fb_pix = s->fb_pix;
  CU(cudaEventRecord(s->ev[0], st));
  CU(cudaMemsetAsync(&s->p2p.d_mail->pc, 0, GVT_PERSIST_RESET_BYTES, st));
  CU(cudaMemsetAsync(s->d_ctr, 0, sizeof(DevCounters), st));
  if (me == root) CU(cudaMemsetAsync(s->d_fb, 0, s->fb_pix * 16, st)); // clearBuffer (TracerBase.h:310); the peers add into it
  s->n_launch_rec = 0;
  LAUNCHED(1);
  if (count)
    k_persist<true><<<grid, TRACE_THREADS, 0, st>>>(fp, scene_dev(*s), pp, g_tune);
  else
    k_persist<false><<<grid, TRACE_THREADS, 0, st>>>(fp, scene_dev(*s), pp, g_tune);
  CU(cudaEventRecord(s->ev[1], st));
  for (int k = 2; k < 5; ++k) CU(cudaEventRecord(s->ev[k], st));
  CU(cudaGetLastError());
  s->timing_valid = true;
  return GVTB200_OK;
}
int gvtb200_domain_persist_profile(gvtb200_scene s, uint64_t prof[16]) {
  NEED_SCENE();
  if (!s->p2p.d_mail || !prof) return fail(GVTB200_ERR_INVALID, "no persistent frame has run / NULL argument");
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaMemcpy(prof, s->p2p.d_mail->pc.prof, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return GVTB200_OK;
}
int gvtb200_domain_persist_progress(gvtb200_scene s, uint32_t *samples, int cap, int *n) {
  NEED_SCENE();
  if (!s->p2p.d_mail || !samples || !n) return fail(GVTB200_ERR_INVALID, "no persistent frame has run / NULL argument");
  CU(cudaStreamSynchronize(s->stream));
  PersistCtl *pc = &s->p2p.d_mail->pc;
  unsigned k = 0;
  CU(cudaMemcpy(&k, &pc->n_progress, 4, cudaMemcpyDeviceToHost));
  k = std::min<unsigned>(std::min<unsigned>(k, 256u), (unsigned)std::max(0, cap));
  if (k) CU(cudaMemcpy(samples, pc->progress, (size_t)k * 12, cudaMemcpyDeviceToHost));
  *n = (int)k;
  return GVTB200_OK;
}
int gvtb200_domain_persist_stats(gvtb200_scene s, double *busy_ms, uint64_t *inbox_rays, uint64_t *instance_load, int n_instances) {
  NEED_SCENE();
  if (!s->p2p.d_mail) return fail(GVTB200_ERR_INVALID, "no persistent frame has run");
  CU(cudaStreamSynchronize(s->stream));
  PersistCtl pc;
  CU(cudaMemcpy(&pc, &s->p2p.d_mail->pc, sizeof(pc), cudaMemcpyDeviceToHost));
  if (busy_ms) *busy_ms = pc.t_last > pc.t_go ? (double)(pc.t_last - pc.t_go) * 1e-6 : 0.0;
  if (inbox_rays) *inbox_rays = pc.tail;
  if (instance_load) {
    if (n_instances != (int)s->inst.size() || !s->p2p.d_inst_load) return fail(GVTB200_ERR_INVALID, "instance_load needs %zu entries", s->inst.size());
    CU(cudaMemcpy(instance_load, s->p2p.d_inst_load, (size_t)n_instances * 8, cudaMemcpyDeviceToHost));
  }
  return GVTB200_OK;
}

// ---- BLAS sharing between ranks: a replica in spare HBM, or the owner's arrays mapped over NVLink ------------------------
int gvtb200_mesh_export(gvtb200_mesh m, gvtb200_mesh_desc *out) {
  NEED_INIT();
  if (!m || !out) return fail(GVTB200_ERR_INVALID, "NULL argument");
  if (m->imported) return fail(GVTB200_ERR_INVALID, "an imported mesh cannot be exported again: export the owner's");
  CU(cudaSetDevice(m->device));
  memset(out, 0, sizeof(*out));
  const void *ptrs[6] = { m->d_nodes, m->d_tris, m->d_faces, m->d_normals, m->d_vcolors, m->d_mats };
  const size_t bytes[6] = { GVT_NODE_WORDS * sizeof(float4) * std::max<size_t>(1, m->stats[0]), 3 * sizeof(float4) * std::max<size_t>(1, m->tri_slots),
                            sizeof(int4) * std::max<size_t>(1, m->nf), sizeof(float4) * std::max<size_t>(1, m->nv), sizeof(float4) * std::max<size_t>(1, m->nv),
                            2 * sizeof(float4) * (size_t)m->n_mats };
  for (int k = 0; k < 6; ++k) {
    if (!ptrs[k]) continue;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, (void *)ptrs[k]));
    memcpy(out->handle[k], &h, 64);
    out->raw[k] = (uint64_t)(uintptr_t)ptrs[k], out->bytes[k] = bytes[k];
  }
  memcpy(out->mat, &m->blas.mat0, 32);
  out->n_tris = m->nf, out->n_nodes = (uint32_t)m->stats[0], out->n_vertices = m->nv, out->n_mats = m->n_mats;
  out->tri_slots = m->tri_slots;
  out->pid = (int32_t)getpid(), out->device = m->device;
  return GVTB200_OK;
}
int gvtb200_mesh_import(const gvtb200_mesh_desc *d, int mode, gvtb200_mesh *out) {
  NEED_INIT();
  if (!d || !out) return fail(GVTB200_ERR_INVALID, "NULL argument");
  if (mode != GVTB200_IMPORT_COPY && mode != GVTB200_IMPORT_MAP) return fail(GVTB200_ERR_INVALID, "mode %d", mode);
  *out = nullptr;
  const int dev = cur_device();
  CU(cudaSetDevice(dev));
  gvtb200_mesh m = new gvtb200_mesh_s();
  memset(&m->blas, 0, sizeof(Blas));
  m->device = dev, m->imported = true;
  m->nv = d->n_vertices, m->nf = d->n_tris, m->n_mats = d->n_mats, m->tri_slots = d->tri_slots;
  m->stats[0] = d->n_nodes, m->stats[1] = d->n_tris;
  void *local[6] = {};
  const bool same_process = d->pid == (int32_t)getpid();
  int rc = GVTB200_OK;
  for (int k = 0; k < 6 && rc == GVTB200_OK; ++k) {
    if (!d->raw[k]) continue;
    void *src = nullptr;
    if (same_process) {
      if ((rc = enable_peer(dev, d->device))) break;
      src = (void *)(uintptr_t)d->raw[k];
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, d->handle[k], 64);
      cudaError_t e = cudaIpcOpenMemHandle(&src, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        rc = fail(GVTB200_ERR_CUDA, "cudaIpcOpenMemHandle(mesh array %d): %s", k, cudaGetErrorString(e));
        break;
      }
    }
    if (mode == GVTB200_IMPORT_COPY) {
      cudaError_t e = cudaMalloc(&local[k], std::max<size_t>(16, d->bytes[k]));
      if (e == cudaSuccess) e = cudaMemcpy(local[k], src, d->bytes[k], cudaMemcpyDefault);
      if (!same_process) cudaIpcCloseMemHandle(src);
      if (e != cudaSuccess) rc = fail(e == cudaErrorMemoryAllocation ? GVTB200_ERR_OOM : GVTB200_ERR_CUDA, "replica of mesh array %d: %s", k, cudaGetErrorString(e));
    } else {
      local[k] = src;
      if (same_process) m->borrowed[k] = true; else m->ipc_opened[k] = src;
    }
  }
  m->d_nodes = (float4 *)local[0], m->d_tris = (float4 *)local[1], m->d_faces = (int4 *)local[2], m->d_normals = (float4 *)local[3];
  m->d_vcolors = (float4 *)local[4], m->d_mats = (float4 *)local[5];
  if (rc != GVTB200_OK) {
    gvtb200_mesh_destroy(m);
    return rc;
  }
  memcpy(&m->blas.mat0, d->mat, 32);
  m->blas.nodes = m->d_nodes, m->blas.tris = m->d_tris, m->blas.faces = m->d_faces, m->blas.normals = m->d_normals;
  m->blas.vcolors = m->d_vcolors, m->blas.mats = m->d_mats;
  m->blas.n_tris = d->n_tris, m->blas.n_nodes = d->n_nodes;
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    gvtb200_mesh_destroy(m);
    return fail(GVTB200_ERR_CUDA, "stream: %s", cudaGetErrorString(e));
  }
  *out = m;
  return GVTB200_OK;
}

int gvtb200_domain_composite_p2p(gvtb200_scene s, int root) {
  NEED_SCENE();
  if (!s->p2p.attached) return fail(GVTB200_ERR_INVALID, "scene is not attached to its peers (gvtb200_p2p_attach)");
  if (root < 0 || root >= s->p2p.world) return fail(GVTB200_ERR_INVALID, "root %d outside 0..%d", root, s->p2p.world - 1);
  cudaStream_t st = s->stream;
  const unsigned seq = ++s->p2p.seq;
  Mailbox *mine = s->p2p.d_mail;
  if (s->p2p.rank == root) {
    k_p2p_wait<<<1, GVT_P2P_MAX, 0, st>>>(mine->done, s->p2p.world, root, seq, &mine->err, g_peer_timeout_ns);
    k_fb_clamp<<<blocks_for(s->fb_pix, 256), 256, 0, st>>>(s->d_fb, s->fb_pix);
  } else {
    k_fb_accumulate_root<<<g_dev[s->device].sms * 8, 256, 0, st>>>((const float4 *)s->d_fb, s->p2p.peer_fb[root], s->fb_pix, mine, s->p2p.mail.m[root], s->p2p.rank, seq);
  }
  LAUNCHED(s->p2p.rank == root ? 2 : 1);
  CU(cudaGetLastError());
  return GVTB200_OK;
}

} // extern "C\"
