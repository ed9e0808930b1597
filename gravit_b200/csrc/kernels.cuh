// kernels.cuh -- __global__ kernels for sm_100a: BLAS build (Morton, radix sort, LBVH, wide collapse), camera +
// routing, the persistent two-level trace kernel, framebuffer and domain-exchange helpers.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cuda_runtime.h>

#include "bvh_build.h"
#include "gvt_device.h"
#include "shade.h"
#include "traverse.h"

namespace cg = cooperative_groups;

namespace gvt {

// =========================================================================================================
// BLAS build
// =========================================================================================================
__global__ void k_tri_setup(BuildCtx c) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.n) tri_setup(c, i);
}
__global__ void k_leaf_setup(BuildCtx c) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.n) leaf_setup(c, i);
}
__global__ void k_karras(BuildCtx c) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < c.n) karras_node(c, (int)i);
}
// bottom-up refit: the second thread to reach an internal node computes its box and continues upwards
__global__ void k_refit(BuildCtx c) {
  unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n) return;
  int node = c.bn_parent[c.n - 1 + k];
  if (c.n == 1) return;
  while (node >= 0) {
    __threadfence();
    unsigned old = atomicAdd(&c.bn_flag[node], 1u);
    if (old == 0u) return;
    int2 ch = c.bn_child[node];
    // children boxes were written by other threads: read through L2
    float4 l0 = __ldcg(&c.bn_lo[ch.x]), h0 = __ldcg(&c.bn_hi[ch.x]);
    float4 l1 = __ldcg(&c.bn_lo[ch.y]), h1 = __ldcg(&c.bn_hi[ch.y]);
    c.bn_lo[node] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
    c.bn_hi[node] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
    node = c.bn_parent[node];
  }
}
// ---- insertion-based optimisation of the binary tree (bodies and the algorithm: bvh_build.h reinsert_*) ------------------
// one thread per binary node (internal nodes 0..n-2, leaves n-1..2n-2; node 0 is the root and never moves)
__global__ void k_ri_search(BuildCtx c) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x == 0) c.ri_pick[0] = make_int4(-1, -1, 0, 0);
  if (x >= 1 && x < 2 * c.n - 1) reinsert_search(c, (int)x);
}
__global__ void k_ri_lock(BuildCtx c) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= 1 && x < 2 * c.n - 1) reinsert_lock(c, (int)x);
}
__global__ void k_ri_check(BuildCtx c) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= 1 && x < 2 * c.n - 1) reinsert_check(c, (int)x);
}
__global__ void k_ri_confirm(BuildCtx c) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x == 0) c.ri_go[0] = 0;
  if (x >= 1 && x < 2 * c.n - 1) reinsert_confirm(c, (int)x);
}
__global__ void k_ri_apply(BuildCtx c) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= 1 && x < 2 * c.n - 1) reinsert_apply(c, (int)x);
}
// triangles below every internal node, bottom-up (same climbing rule as k_refit)
__global__ void k_ri_count(BuildCtx c) {
  unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n || c.n < 2) return;
  int node = c.bn_parent[c.n - 1 + k];
  while (node >= 0) {
    __threadfence();
    unsigned old = atomicAdd(&c.bn_flag[node], 1u);
    if (old == 0u) return;
    relayout_count_node(c, node);
    node = c.bn_parent[node];
  }
}
// relayout, step 1 (reads only): new sorted position, parent and child slot of every leaf; ranges of the internal nodes
__global__ void k_ri_layout_a(BuildCtx c, unsigned *newpos, int *leaf_parent, int *leaf_slot) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c.n) {
    const int id = (int)(c.n - 1 + i), p = c.bn_parent[id];
    newpos[i] = relayout_first(c, id), leaf_parent[i] = p, leaf_slot[i] = c.bn_child[p].y == id ? 1 : 0;
  } else if (i < 2 * c.n - 1) {
    const int b = (int)(i - c.n);
    const unsigned f = relayout_first(c, b);
    c.bn_range[b] = make_int2((int)f, (int)(f + c.ri_cnt[b] - 1u));
  }
}
// step 2: leaves take their new numbers (n - 1 + position); the element order follows
__global__ void k_ri_layout_b(BuildCtx c, const unsigned *newpos, const int *leaf_parent, const int *leaf_slot, uint32_t *vals_out) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n) return;
  const unsigned np = newpos[k];
  const int nid = (int)(c.n - 1 + np), p = leaf_parent[k];
  vals_out[np] = c.vals[k];
  c.bn_parent[nid] = p;
  if (leaf_slot[k]) c.bn_child[p].y = nid; else c.bn_child[p].x = nid;
}
// optimal-widening tables, bottom-up: the second thread to reach an internal node fills its row (both children's rows
// are complete by then) and continues upwards
__global__ void k_collapse_dp(BuildCtx c) {
  unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n || c.n < 2) return;
  int node = c.bn_parent[c.n - 1 + k];
  while (node >= 0) {
    __threadfence();
    unsigned old = atomicAdd(&c.bn_flag[node], 1u);
    if (old == 0u) return;
    collapse_dp_node(c, node);
    node = c.bn_parent[node];
  }
}
// one thread per wide node of the current level; children tasks are appended to `out`
__global__ void k_collapse(BuildCtx c, const CollapseTask *in, unsigned n_in, CollapseTask *out, unsigned *n_out) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_in) return;
  CollapseTask t = in[i];
  WideNode w;
  collapse_node(c, t, w);
  unsigned child_base = w.n_internal ? atomicAdd(c.node_count, (unsigned)w.n_internal) : 0u;
  unsigned tri_base = w.n_leaves ? atomicAdd(c.tri_count, (unsigned)(GVT_LEAF_MAX * w.n_leaves)) : 0u;
  unsigned qpos = w.n_internal ? atomicAdd(n_out, (unsigned)w.n_internal) : 0u;
  if (w.n_tris) atomicAdd(c.tri_real, (unsigned)w.n_tris);
  collapse_emit(c, t, w, child_base, tri_base, out + qpos);
}

// ---- binned-SAH tree, one level per round (bodies in bvh_build.h) ------------------------------------------------
struct SahCtx {
  BuildCtx c;
  float4 *tb_lo, *tb_hi;     // padded triangle boxes, by triangle id
  uint32_t *idx, *idx2;      // element order (triangle ids), current / next level
  int *node_of, *node_of2;   // active-node index per position, -1 = settled (leaf)
  SahNode *act, *act_next;   // active nodes of this / the next level
  SahBin *bins;              // 3 * GVT_SAH_BINS per big active node
  unsigned *flag;            // n + 1 goes-left flags -> exclusive scan
  unsigned *counters;        // [0] next internal node id, [1] active nodes of the next level, [2] bin slots of the next level
};
__global__ void k_sah_tbox(SahCtx s) {
  unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= s.c.n) return;
  float4 lo, hi;
  tri_box_padded(s.c, f, lo, hi);
  s.tb_lo[f] = lo, s.tb_hi[f] = hi;
}
__global__ void k_sah_reset(SahCtx s, unsigned m, unsigned n_bins) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) {
    for (int q = 0; q < 3; ++q) s.act[i].cb[q] = 0x7FFFFFFF, s.act[i].cb[3 + q] = (int)0x80000000;
  }
  if (i < n_bins) {
    SahBin b;
    for (int q = 0; q < 3; ++q) b.lo[q] = 0x7FFFFFFF, b.hi[q] = (int)0x80000000;
    b.cnt = 0, b.pad = 0;
    s.bins[i] = b;
  }
}
// The elements of an active node are contiguous, so at the top of the tree whole blocks belong to one node and every
// thread would hit the same few words with its atomics (measured: 3.8 ns per same-address atomic, 0.28 ms per level for
// 75 K triangles, 28 ms per level for 7 M).  A block that lies inside one node reduces in shared memory first and issues one
// global atomic per word; a warp that lies inside one node reduces with warp votes; only mixed warps fall back to per-lane
// atomics.  Integer min / max / add: the result does not depend on the order.
__device__ __forceinline__ int sah_block_node(const SahCtx &s) { // the block's node when all its positions share it, else -1
  const unsigned first = blockIdx.x * blockDim.x;
  if (first >= s.c.n) return -1;
  const unsigned last = min(first + blockDim.x, s.c.n) - 1u;
  const int a0 = s.node_of[first], a1 = s.node_of[last];
  return (a0 == a1) ? a0 : -1; // node ranges are disjoint intervals: equal ends mean equal in between
}
__global__ void k_sah_cb(SahCtx s) {
  __shared__ int sm[6];
  const unsigned pos = blockIdx.x * blockDim.x + threadIdx.x;
  const int ab = sah_block_node(s); // block-uniform
  if (threadIdx.x < 6) sm[threadIdx.x] = threadIdx.x < 3 ? 0x7FFFFFFF : (int)0x80000000;
  __syncthreads();
  const bool valid = pos < s.c.n;
  const int a = valid ? s.node_of[pos] : -1;
  int e[3] = { 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF }, g[3] = { (int)0x80000000, (int)0x80000000, (int)0x80000000 };
  if (a >= 0) {
    const uint32_t f = s.idx[pos];
    const float4 l = s.tb_lo[f], h = s.tb_hi[f];
    e[0] = g[0] = fenc(0.5f * (l.x + h.x)), e[1] = g[1] = fenc(0.5f * (l.y + h.y)), e[2] = g[2] = fenc(0.5f * (l.z + h.z));
  }
  const unsigned FULL = 0xffffffffu;
  const int a_first = __shfl_sync(FULL, a, 0);
  const bool warp_uniform = __all_sync(FULL, a == a_first || !valid) && a_first >= 0;
  if (ab >= 0 || warp_uniform) {
    for (int q = 0; q < 3; ++q) e[q] = __reduce_min_sync(FULL, e[q]), g[q] = __reduce_max_sync(FULL, g[q]);
    if ((threadIdx.x & 31u) == 0u) {
      if (ab >= 0) {
        for (int q = 0; q < 3; ++q) atomicMin(&sm[q], e[q]), atomicMax(&sm[3 + q], g[q]);
      } else {
        for (int q = 0; q < 3; ++q) atomicMin(&s.act[a_first].cb[q], e[q]), atomicMax(&s.act[a_first].cb[3 + q], g[q]);
      }
    }
  } else if (a >= 0) {
    for (int q = 0; q < 3; ++q) atomicMin(&s.act[a].cb[q], e[q]), atomicMax(&s.act[a].cb[3 + q], g[q]);
  }
  if (ab >= 0) {
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&s.act[ab].cb[threadIdx.x], sm[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&s.act[ab].cb[threadIdx.x], sm[threadIdx.x]);
  }
}
__global__ void k_sah_bin(SahCtx s) {
  __shared__ int sb[3 * GVT_SAH_BINS * 7]; // per (axis, bin): lo.xyz, hi.xyz, count
  const unsigned pos = blockIdx.x * blockDim.x + threadIdx.x;
  const int ab = sah_block_node(s);
  const bool shared_bins = ab >= 0 && s.act[ab].bslot >= 0;
  if (shared_bins) {
    for (unsigned k = threadIdx.x; k < 3u * GVT_SAH_BINS * 7u; k += blockDim.x) {
      const unsigned w = k % 7u;
      sb[k] = w < 3u ? 0x7FFFFFFF : (w < 6u ? (int)0x80000000 : 0);
    }
    __syncthreads();
  }
  if (pos < s.c.n) {
    const int a = s.node_of[pos];
    if (a >= 0) {
      const SahNode &nd = s.act[a];
      if (nd.bslot >= 0) {
        const uint32_t f = s.idx[pos];
        const float4 l = s.tb_lo[f], h = s.tb_hi[f];
        const int el[3] = { fenc(l.x), fenc(l.y), fenc(l.z) }, eh[3] = { fenc(h.x), fenc(h.y), fenc(h.z) };
        for (int ax = 0; ax < 3; ++ax) {
          const float c = 0.5f * ((&l.x)[ax] + (&h.x)[ax]);
          const int b = sah_bin(c, fdec(nd.cb[ax]), fdec(nd.cb[3 + ax]));
          if (shared_bins) {
            int *B = sb + (ax * GVT_SAH_BINS + b) * 7;
            for (int q = 0; q < 3; ++q) atomicMin(&B[q], el[q]), atomicMax(&B[3 + q], eh[q]);
            atomicAdd(&B[6], 1);
          } else {
            SahBin *B = s.bins + ((size_t)nd.bslot * 3 + ax) * GVT_SAH_BINS + b;
            for (int q = 0; q < 3; ++q) atomicMin(&B->lo[q], el[q]), atomicMax(&B->hi[q], eh[q]);
            atomicAdd(&B->cnt, 1);
          }
        }
      }
    }
  }
  if (shared_bins) {
    __syncthreads();
    const SahNode &nd = s.act[ab];
    for (unsigned k = threadIdx.x; k < 3u * GVT_SAH_BINS * 7u; k += blockDim.x) {
      const unsigned bin = k / 7u, w = k % 7u;
      if (sb[bin * 7u + 6u] == 0) continue; // nothing fell into this bin
      SahBin *B = s.bins + (size_t)nd.bslot * 3 * GVT_SAH_BINS + bin;
      if (w < 3u) atomicMin(&B->lo[w], sb[k]);
      else if (w < 6u) atomicMax(&B->hi[w - 3u], sb[k]);
      else atomicAdd(&B->cnt, sb[k]);
    }
  }
}
__global__ void k_sah_split(SahCtx s, unsigned m) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  SahNode nd = s.act[i];
  float lo[3], hi[3];
  if (nd.bslot >= 0)
    sah_eval_bins(s.bins + (size_t)nd.bslot * 3 * GVT_SAH_BINS, nd, lo, hi);
  else
    sah_eval_small(s.tb_lo, s.tb_hi, s.idx, nd, lo, hi);
  s.c.bn_lo[nd.id] = make_float4(lo[0], lo[1], lo[2], 0.f);
  s.c.bn_hi[nd.id] = make_float4(hi[0], hi[1], hi[2], 0.f);
  s.act[i] = nd;
}
__global__ void k_sah_flag(SahCtx s) {
  unsigned pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos > s.c.n) return;
  unsigned fl = 0;
  if (pos < s.c.n) {
    int a = s.node_of[pos];
    if (a >= 0) fl = sah_goes_left(s.act[a], s.tb_lo, s.tb_hi, s.idx[pos], (int)pos) ? 1u : 0u;
  }
  s.flag[pos] = fl;
}
__global__ void k_sah_children(SahCtx s, unsigned m) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  SahNode nd = s.act[i];
  const int n = (int)s.c.n, cnt = nd.hi - nd.lo;
  int nl = (int)(s.flag[nd.hi] - s.flag[nd.lo]);
  if (nl <= 0 || nl >= cnt) nd.mode = 1, nl = cnt / 2; // cannot happen with consistent binning; keep the tree valid anyway
  nd.n_left = nl;
  int child[2];
  for (int side = 0; side < 2; ++side) {
    int lo = side ? nd.lo + nl : nd.lo, hi = side ? nd.hi : nd.lo + nl;
    if (hi - lo == 1) {
      child[side] = n - 1 + lo;
      nd.child_active[side] = -1;
    } else {
      int id = (int)atomicAdd(&s.counters[0], 1u);
      unsigned ai = atomicAdd(&s.counters[1], 1u);
      SahNode ch;
      ch.lo = lo, ch.hi = hi, ch.id = id;
      ch.bslot = (hi - lo > GVT_SAH_SMALL) ? (int)atomicAdd(&s.counters[2], 1u) : -1;
      ch.axis = ch.bin = ch.mode = ch.n_left = 0, ch.thr = 0.f, ch.child_active[0] = ch.child_active[1] = -1;
      for (int q = 0; q < 6; ++q) ch.cb[q] = 0;
      s.act_next[ai] = ch;
      s.c.bn_range[id] = make_int2(lo, hi - 1);
      child[side] = id;
      nd.child_active[side] = (int)ai;
    }
    s.c.bn_parent[child[side]] = nd.id;
  }
  s.c.bn_child[nd.id] = make_int2(child[0], child[1]);
  s.act[i] = nd;
}
__global__ void k_sah_scatter(SahCtx s) {
  unsigned pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= s.c.n) return;
  int a = s.node_of[pos];
  uint32_t f = s.idx[pos];
  if (a < 0) {
    s.idx2[pos] = f, s.node_of2[pos] = -1;
    return;
  }
  const SahNode &nd = s.act[a];
  unsigned np;
  bool left;
  if (nd.mode == 1) {
    left = ((int)pos - nd.lo) < nd.n_left;
    np = pos;
  } else {
    left = (s.flag[pos + 1] - s.flag[pos]) != 0u;
    unsigned lbefore = s.flag[pos] - s.flag[nd.lo];
    np = left ? (unsigned)nd.lo + lbefore : (unsigned)(nd.lo + nd.n_left) + (pos - (unsigned)nd.lo) - lbefore;
  }
  s.idx2[np] = f;
  s.node_of2[np] = nd.child_active[left ? 0 : 1];
}

// ---- LSD radix sort, 8 bits per pass, stable; keys + values (uint32) ---------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)

__global__ void k_rs_hist(const uint32_t *keys, unsigned n, int shift, unsigned *hist, unsigned nblocks) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  unsigned base = blockIdx.x * RS_TILE;
  for (int it = 0; it < RS_ITEMS; ++it) {
    unsigned i = base + it * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// in-place exclusive scan of one tile per block; block totals to sums[]
__global__ void k_scan_tile(unsigned *data, unsigned n, unsigned *sums) {
  __shared__ unsigned warp_tot[RS_THREADS / 32];
  unsigned base = blockIdx.x * RS_TILE + threadIdx.x * RS_ITEMS;
  unsigned v[RS_ITEMS], tot = 0;
  for (int k = 0; k < RS_ITEMS; ++k) {
    v[k] = (base + k < n) ? data[base + k] : 0u;
    tot += v[k];
  }
  unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = tot;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < RS_THREADS / 32 ? warp_tot[lane] : 0u;
    unsigned winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (unsigned)o) winc += t;
    }
    if (lane < RS_THREADS / 32) warp_tot[lane] = winc - w;
    if (lane == RS_THREADS / 32 - 1 && sums) sums[blockIdx.x] = winc;
  }
  __syncthreads();
  unsigned run = warp_tot[warp] + inc - tot;
  for (int k = 0; k < RS_ITEMS; ++k) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
}
__global__ void k_scan_add(unsigned *data, unsigned n, const unsigned *sums) {
  unsigned base = blockIdx.x * RS_TILE;
  unsigned add = sums[blockIdx.x];
  for (int it = 0; it < RS_ITEMS; ++it) {
    unsigned i = base + it * RS_THREADS + threadIdx.x;
    if (i < n) data[i] += add;
  }
}

__global__ void k_rs_scatter(const uint32_t *keys, const uint32_t *vals, uint32_t *okeys, uint32_t *ovals, unsigned n, int shift,
                             const unsigned *offs, unsigned nblocks) {
  __shared__ unsigned cnt[RS_THREADS / 32][256];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = 0; w < RS_THREADS / 32; ++w) cnt[w][threadIdx.x] = 0;
  __syncthreads();
  // warp `warp` owns RS_TILE/8 consecutive keys, walked 32 at a time in order (stability)
  const unsigned wbase = blockIdx.x * RS_TILE + warp * (RS_TILE / (RS_THREADS / 32));
  uint32_t k[RS_ITEMS], v[RS_ITEMS];
  const unsigned ltmask = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    unsigned i = wbase + it * 32 + lane;
    bool ok = i < n;
    k[it] = ok ? keys[i] : 0xFFFFFFFFu;
    v[it] = ok ? vals[i] : 0u;
    unsigned d = ok ? ((k[it] >> shift) & 255u) : 256u + lane; // invalid lanes get private "digits"
    unsigned peers = __match_any_sync(0xffffffffu, d);
    if (ok && (peers & ltmask) == 0u) cnt[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  { // digit threadIdx.x: exclusive prefix over the warps, seeded with this block's global offset
    unsigned d = threadIdx.x;
    unsigned run = offs[d * nblocks + blockIdx.x];
    for (int w = 0; w < RS_THREADS / 32; ++w) {
      unsigned c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    unsigned i = wbase + it * 32 + lane;
    bool ok = i < n;
    unsigned d = ok ? ((k[it] >> shift) & 255u) : 256u + lane;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    if (ok) {
      unsigned pos = cnt[warp][d] + __popc(peers & ltmask);
      okeys[pos] = k[it];
      ovals[pos] = v[it];
    }
    __syncwarp();
    if (ok && (peers & ltmask) == 0u) cnt[warp][d] += __popc(peers);
    __syncwarp();
  }
}

// =========================================================================================================
// Frame kernels
// =========================================================================================================
struct SceneDev {
  Tlas tlas;
  const Instance *inst;
  const Blas *blas;
  float *fb; // RGBA float, W*H*4
  DevCounters *ctr;
};

// warp-aggregated append: one atomic per group of converged lanes
__device__ __forceinline__ unsigned queue_reserve(unsigned *count) {
  cg::coalesced_group g = cg::coalesced_threads();
  unsigned base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(count, g.size());
  return g.shfl(base, 0) + g.thread_rank();
}
// the same for a queue that lives in another GPU's memory: lanes are grouped by destination rank, the group leader does
// one system-scope atomic over NVLink
__device__ __forceinline__ unsigned queue_reserve_peer(unsigned *count, int owner) {
  cg::coalesced_group g = cg::labeled_partition(cg::coalesced_threads(), owner);
  unsigned base = 0;
  if (g.thread_rank() == 0) base = atomicAdd_system(count, g.size());
  return g.shfl(base, 0) + g.thread_rank();
}
__device__ __forceinline__ void count_add(unsigned long long *c, unsigned long long v) {
  cg::coalesced_group g = cg::coalesced_threads();
  unsigned long long s = cg::reduce(g, v, cg::plus<unsigned long long>());
  if (g.thread_rank() == 0 && s) atomicAdd(c, s);
}

// IceTComposite::localAdd (composite/IceTComposite.cpp:111-117): += colour, alpha += 1.  The running clamp to 1 is
// applied once at the end (k_fb_clamp): min(1, sum) == clamp-after-each-add for non-negative terms.
__device__ __forceinline__ void fb_add(float *fb, int id, V3 c) {
  atomicAdd(reinterpret_cast<float4 *>(fb) + id, make_float4(c.x, c.y, c.z, 1.0f)); // one 16-byte RED (sm_90+)
}
__global__ void k_fb_clamp(float *fb, size_t npix) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  float4 v = reinterpret_cast<float4 *>(fb)[i];
  v.x = fminf(v.x, 1.f), v.y = fminf(v.y, 1.f), v.z = fminf(v.z, 1.f);
  reinterpret_cast<float4 *>(fb)[i] = v;
}

// Which film rows a frame call covers (ImageScheduler sharding, ImageTracer.h:111-125): rows
//   row0 + (b * stride + index) * tile_rows + k,  k < tile_rows, b = 0,1,...   clipped to [row0, row1).
// stride == 1 gives the contiguous block [row0, row1).
struct RowTiles {
  int row0, row1, tile_rows, index, stride;
};
__host__ __device__ inline size_t tiles_local_rows(const RowTiles &t) { // rows this shard owns
  int span = t.row1 - t.row0;
  if (span <= 0) return 0;
  int period = t.tile_rows * t.stride;
  int full = span / period, rem = span - full * period;
  int extra = rem - t.index * t.tile_rows;
  extra = extra < 0 ? 0 : (extra > t.tile_rows ? t.tile_rows : extra);
  return (size_t)full * t.tile_rows + extra;
}

// camera rays of this shard's rows -> initial routing (FilterRaysLocally: TracerBase.h:319 / DomainTracer.h:148-181)
__global__ void k_camera_route(FrameParams fp, SceneDev sd, RowTiles tiles, size_t count, RayQueue out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long routed = 0, qovf = 0;
  const size_t per_row = (size_t)fp.cam.W * fp.cam.samples * fp.cam.samples;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += stride) {
    size_t lr = i / per_row, within = i - lr * per_row;
    size_t b = lr / tiles.tile_rows, k = lr - b * tiles.tile_rows;
    size_t row = (size_t)tiles.row0 + (b * tiles.stride + tiles.index) * tiles.tile_rows + k;
    Ray r;
    camera_ray(fp.cam, row * per_row + within, r);
    float t;
    int next = tlas_route(sd.tlas, r.o, r.d, r.t_max, -1, (fp.flags & FLAG_ENTER_INSIDE) != 0, t);
    if (next < 0) continue;
    if (fp.world > 1 && sd.inst[next].owner != fp.rank) continue; // shuffleDropRays keeps local first hits only
    r.o = xadd(r.o, xmul(r.d, xmul(t, 0.95f)));                   // TracerBase.h:394
    unsigned pos = queue_reserve(out.count);
    if (pos < out.cap) queue_store(out, pos, r, next);
    else qovf++;
    routed++;
  }
  count_add(&sd.ctr->routed_initial, routed);
  count_add(&sd.ctr->queue_overflows, qovf);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&sd.ctr->camera_rays, (unsigned long long)count);
}

// camera rays only (parity hook for gvtPerspectiveCamera::generateRays)
__global__ void k_camera_only(FrameParams fp, size_t count, RayQueue out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  Ray r;
  camera_ray(fp.cam, i, r);
  queue_store(out, (unsigned)i, r, -1);
}

// routing only (parity hook for BVH::intersect / shuffleRays)
__global__ void k_route(FrameParams fp, SceneDev sd, RayQueue in, unsigned n, int from, int *next, float *tout) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ray r;
  int dummy;
  queue_load(in, i, r, dummy);
  float t;
  next[i] = tlas_route(sd.tlas, r.o, r.d, r.t_max, from, from == -1 && (fp.flags & FLAG_ENTER_INSIDE) != 0, t);
  tout[i] = t;
}

// ---------------------------------------------------------------------------------------------------------
// The hot kernels: k_trace (traversal + routing) and k_shade (shading + ray generation).  Together they are
// Adapter::trace + shuffleRays (EmbreeMeshAdapter.cpp:436-623 + TracerBase.h:325-414).
//
// k_trace: persistent CTAs.  Every lane owns one ray at a time and runs it to the end of its current life: trace in
// its instance (closest hit, or any hit for shadow rays); on a miss, route to the next instance (from = this one),
// advance the origin and carry on; a closest hit is appended -- ray + hit record -- to the hit queue `hq` for k_shade;
// an occluded shadow ray is dropped; a shadow ray that leaves the scene adds its colour to the framebuffer.
// The warp runs in lock step over "iterations" (one node visit, then triangle tests) so that
//   * lanes whose ray is finished are refilled from the input queue as soon as `refill_at` lanes are waiting
//     (one atomic per refill), instead of idling until the slowest ray of a 32-ray bundle ends, and
//   * triangle tests are postponed (the triangle group goes onto the traversal stack) while fewer than
//     postpone_num/8 of the traversing lanes have a triangle to test,
// after Ylitie, Karras, Laine (HPG 2017) sections 4.3-4.4.  The traversal stack lives in shared memory
// (TRACE_SM_STACK entries per lane, conflict-free 8-byte accesses) and spills to local memory beyond that.
//   mode 0: scheduler mode (two-level, as above)
//   mode 1: adapter mode (Adapter::trace contract): a ray that misses is appended to `mvq` instead of being routed
// DomainScheduler: a ray routed to an instance owned by another rank is parked in `exq` (inst column = instance).
// ---------------------------------------------------------------------------------------------------------
#define TRACE_THREADS 128
#ifndef TRACE_SM_STACK
#define TRACE_SM_STACK 12
#endif
#ifndef TRACE_MIN_BLOCKS
#define TRACE_MIN_BLOCKS 6 /* CTAs per SM the register allocation is bounded for */
#endif

struct TraceTune {
  int refill_at;    // refill when at least this many lanes are idle (1..32)
  int exit_at;      // leave the traversal loop when at least this many lanes are not traversing and work is waiting
  int postpone_num; // postpone triangle tests while lanes_with_triangles * 8 < traversing_lanes * postpone_num
  int node_burst;   // node visits a lane may make per lock-step iteration before the warp votes again (it stops at the
                    // first triangle group)
  int prefetch_below; // lanes: a warp with at most this many traversing lanes prefetches the siblings it pushes (0 = never)
  int split_below;    // persistent kernel: a warp with at most this many traversing lanes shares out the stacks of its rays
                      // among its free lanes (0 = never)
  int persist_trips;  // persistent kernel: lock-step iterations after which a warp looks at its inbox again at the latest
};

struct CameraGen { // in-kernel camera-ray source (replaces reading a queue of camera rays back from HBM)
  int enabled;
  RowTiles tiles;
  unsigned count; // rays of this film shard
};
struct TraceQueues {
  CameraGen gen; // enabled: input ray i is camera ray i of the shard, generated and routed in registers; `in` is unused
  RayQueue in;
  RayQueue hq;    // closest hits: ray (t = hit t) + hit column
  int4 *hit;      // hit column of hq: prim, triangle slot, u bits, v bits
  RayQueue mvq, exq;
  RayQueue fin;   // adapter mode: final state of every input ray, by input slot (Adapter::trace mutates rayList in place)
  unsigned *work; // input cursor
  // NVLink peer path: peer_inbox[p] = rank p's inbound ray queue, mapped into this GPU's address space (CUDA IPC).  When
  // set, a ray bound for another rank's instance is stored there directly instead of being parked in `exq`.
  const RayQueue *peer_inbox;
};
struct ShadeQueues {
  RayQueue hq;
  const int4 *hit;
  RayQueue shq, nxq, fin;
};

struct SmemStack {
  uint2 *s; // entry k of this lane at s[k * TRACE_THREADS]
  uint2 l[GVT_BLAS_STACK - TRACE_SM_STACK];
  bool overflow;
  __device__ __forceinline__ void push(int &sp, uint2 v) {
    if (sp < TRACE_SM_STACK)
      s[sp * TRACE_THREADS] = v;
    else if (sp < GVT_BLAS_STACK)
      l[sp - TRACE_SM_STACK] = v;
    else {
      overflow = true;
      return;
    }
    ++sp;
  }
  __device__ __forceinline__ uint2 pop(int &sp) {
    --sp;
    return sp < TRACE_SM_STACK ? s[sp * TRACE_THREADS] : l[sp - TRACE_SM_STACK];
  }
  __device__ __forceinline__ uint2 get(int k) const { return k < TRACE_SM_STACK ? s[k * TRACE_THREADS] : l[k - TRACE_SM_STACK]; }
  __device__ __forceinline__ void set(int k, uint2 v) {
    if (k < TRACE_SM_STACK) s[k * TRACE_THREADS] = v; else l[k - TRACE_SM_STACK] = v;
  }
};

// The lock-step traversal loop of one warp (shared by k_trace and k_persist): every lane with `active` set advances its
// traversal -- a burst of node visits, then triangle tests, postponed while too few lanes hold a triangle -- until the warp
// should go back to its caller: all lanes are finished, or at least tune.exit_at lanes are not traversing and there is
// something better to do for them (`resting` = lanes that had no ray on entry; a lane outside `resting` that stopped
// traversing has a finished ray to hand over; more_input = the caller can refill).  Returns this lane's `done`.
// BUDGETED: return after `budget` iterations at the latest (the persistent kernel must look at its inbox now and then).
template <bool COUNT, bool BUDGETED, class Stack>
__device__ __forceinline__ bool warp_traverse(bool active, unsigned resting, bool more_input, int budget, const TraceTune &tune, const float4 *nodes,
                                              const float4 *tris, const RayLocal &rl, float tfar, bool any, Trav &tv, Stack &stack,
                                              TravCounters &tc, const unsigned char *s_lut) {
  const unsigned FULL = 0xffffffffu;
  bool done = false;
  int trips = 0;
  for (;;) {
    const bool trav = active && !done;
    const unsigned tm = __ballot_sync(FULL, trav);
    if (tm == 0u) break;
    const unsigned waiting = 32u - (unsigned)__popc(tm);
    // (the persistent kernel always advances one step per call: its callers' lanes can rest in states a refill cannot end)
    if ((!BUDGETED || trips > 0) && waiting >= (unsigned)tune.exit_at && ((~tm & ~resting) != 0u || more_input)) break;
    if (BUDGETED && ++trips > budget) break;
    const bool pf = __popc(tm) <= tune.prefetch_below;
    if (trav) {
#pragma unroll 1
      for (int k = 0; k < tune.node_burst; ++k) {
        if (!trav_has_nodes(tv)) {
          if (k == 0 || trav_has_tris(tv)) break; // first trip: the group popped last time may be a triangle group
          if (!trav_next(tv, stack)) {
            done = true;
            break;
          }
          if (!trav_has_nodes(tv)) break; // popped a postponed triangle group
        }
        node_step<COUNT, true>(nodes, rl, GVT_RAY_EPSILON, tv, stack, &tc, s_lut, pf);
        if (trav_has_tris(tv)) break;
      }
    }
    for (;;) {
      const bool ht = trav && !done && trav_has_tris(tv);
      const unsigned tb = __ballot_sync(FULL, ht);
      if (tb == 0u) break;
      if (__popc(tb) * 8 < __popc(tm) * tune.postpone_num) {
        // too few lanes have a triangle: lanes with node work postpone theirs, the others keep it pending
        if (ht && trav_has_nodes(tv)) {
          stack.push(tv.sp, tv.tg);
          tv.tg.y &= 0xFF000000u;
        }
        break;
      }
      if (ht) done = tri_step<COUNT>(tris, rl, GVT_RAY_EPSILON, tfar, any, tv, &tc);
    }
    if (trav && !done) done = !trav_next(tv, stack);
  }
  return done;
}

template <bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS, TRACE_MIN_BLOCKS) k_trace(FrameParams fp, SceneDev sd, TraceQueues q, int mode, TraceTune tune) {
  __shared__ uint2 s_stack[TRACE_SM_STACK * TRACE_THREADS];
  __shared__ unsigned char s_lut[8 * 256]; // best_slot(hits8, octinv)
  for (unsigned k = threadIdx.x; k < 8u * 256u; k += TRACE_THREADS) s_lut[k] = (unsigned char)best_slot(k & 255u, k >> 8);
  __syncthreads();
  const unsigned FULL = 0xffffffffu;
  const bool gen = q.gen.enabled != 0;
  const unsigned n = gen ? q.gen.count : min(*q.in.count, q.in.cap);
  const unsigned lane = threadIdx.x & 31;
  const unsigned ltmask = (1u << lane) - 1u;
  const bool to_light = (fp.flags & FLAG_SHADOW_TO_LIGHT) != 0;
  unsigned long long c_visits = 0, c_shok = 0, c_fb = 0, c_exp = 0, c_ovf = 0, c_qovf = 0, c_routed = 0;
  const unsigned per_row = (unsigned)(fp.cam.W * fp.cam.samples * fp.cam.samples);
  auto camera_index = [&](unsigned i) -> size_t { // shard ray i -> global camera-ray index (RowTiles arithmetic)
    unsigned lr = i / per_row, within = i - lr * per_row;
    unsigned b = lr / (unsigned)q.gen.tiles.tile_rows, k = lr - b * (unsigned)q.gen.tiles.tile_rows;
    size_t row = (size_t)q.gen.tiles.row0 + ((size_t)b * q.gen.tiles.stride + q.gen.tiles.index) * q.gen.tiles.tile_rows + k;
    return row * per_row + within;
  };
  TravCounters tc;
  tc.nodes = 0, tc.tris = 0;
  SmemStack stack;
  stack.s = s_stack + threadIdx.x;
  stack.overflow = false;

  // the lane's ray
  bool active = false, first = true, any = false;
  bool exhausted = false; // warp-uniform: the input queue has been handed out
  unsigned idx = 0;
  int inst = 0, type = 0;
  V3 o = v3(0.f, 0.f, 0.f), d = o;
  float t_min = 0.f, t_max = 0.f, tfar = FLT_MAX;
  const float4 *nodes = nullptr, *tris = nullptr;
  RayLocal rl;
  Trav tv;
  ray_local(rl, o, o);
  trav_begin(tv, 0u, 0.f);

  auto begin_instance = [&]() {
    const Instance in = sd.inst[inst];
    const Blas &bl = sd.blas[in.mesh];
    V3 lo, ld;
    to_local(in, o, d, lo, ld);
    ray_local(rl, lo, ld);
    nodes = bl.nodes, tris = bl.tris;
    // prepGVT_EMBREE_PACKET_TYPE (EmbreeMeshAdapter.cpp:269-286): tnear = RAY_EPSILON, tfar = FLT_MAX
    tfar = (any && to_light) ? t_max : FLT_MAX;
    trav_begin(tv, bl.n_tris, tfar);
    c_visits++;
  };
  auto reload = [&](Ray &r) { // the columns k_trace does not keep in registers
    if (gen) {
      camera_ray(fp.cam, camera_index(idx), r); // regenerated: cheaper than a round trip through HBM
    } else {
      const float4 c2 = q.in.c2[idx];
      const int4 c3 = q.in.c3[idx];
      r.color = v3(c2.x, c2.y, c2.z), r.t = c2.w;
      r.id = c3.x, r.depth = c3.y, r.w = u2f((uint32_t)c3.z), r.type = c3.w;
    }
    r.o = o, r.t_min = t_min, r.d = d, r.t_max = t_max;
  };

  for (;;) {
    // ---- refill idle lanes from the input queue --------------------------------------------------------
    unsigned idle = __ballot_sync(FULL, !active);
    if (idle != 0u && !exhausted && (idle == FULL || __popc(idle) >= tune.refill_at)) {
      const unsigned cnt = (unsigned)__popc(idle);
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(q.work, cnt);
      base = __shfl_sync(FULL, base, 0);
      if (base + cnt >= n) exhausted = true;
      if (!active) {
        const unsigned my = base + (unsigned)__popc(idle & ltmask);
        if (my < n) {
          idx = my;
          if (gen) { // gvtPerspectiveCamera::generateRays + FilterRaysLocally (first routing hop), in registers
            Ray r;
            camera_ray(fp.cam, camera_index(idx), r);
            float t;
            const int next = tlas_route(sd.tlas, r.o, r.d, r.t_max, -1, (fp.flags & FLAG_ENTER_INSIDE) != 0, t);
            if (next >= 0) {
              c_routed++;
              o = xadd(r.o, xmul(r.d, xmul(t, 0.95f))); // TracerBase.h:394
              t_min = r.t_min, d = r.d, t_max = r.t_max, type = r.type, inst = next;
              any = false, first = true, active = true;
              begin_instance();
            }
          } else {
            const float4 c0 = q.in.c0[idx], c1 = q.in.c1[idx];
            type = q.in.c3[idx].w;
            inst = (mode == 1) ? 0 : q.in.inst[idx];
            o = v3(c0.x, c0.y, c0.z), t_min = c0.w, d = v3(c1.x, c1.y, c1.z), t_max = c1.w;
            any = (type == RAY_SHADOW);
            first = true;
            active = true;
            begin_instance();
          }
        }
      }
      idle = __ballot_sync(FULL, !active);
    }
    if (idle == FULL) {
      if (exhausted) break;
      continue;
    }
    // ---- traverse in lock step ---------------------------------------------------------------------------
    const bool done = warp_traverse<COUNT, false>(active, idle, !exhausted, 0, tune, nodes, tris, rl, tfar, any, tv, stack, tc, s_lut);
    // ---- rays that finished their instance -------------------------------------------------------------------
    if (active && done) {
      if (tv.hit.prim >= 0) {
        if (type != RAY_SHADOW) { // closest hit -> k_shade; an occluded shadow ray is dropped (:487-489)
          Ray r;
          reload(r);
          r.t = tv.hit.t;
          const unsigned pos = queue_reserve(q.hq.count);
          if (pos < q.hq.cap) {
            queue_store(q.hq, pos, r, mode == 1 ? q.in.inst[idx] : inst);
            q.hit[pos] = make_int4(tv.hit.prim, tv.hit.slot, (int)f2u(tv.hit.u), (int)f2u(tv.hit.v));
          } else
            c_qovf++; // the reference never drops a ray (TracerBase.h:403-409): reported as GVTB200_ERR_CAPACITY
        }
        active = false;
      } else {
        if (any && first) c_shok++;
        first = false;
        if (mode == 1) { // adapter contract: hand the ray back (localDispatch, :604-609 / :377-381)
          Ray r;
          reload(r);
          const int slot = q.in.inst[idx];
          const unsigned pos = queue_reserve(q.mvq.count);
          if (pos < q.mvq.cap) queue_store(q.mvq, pos, r, slot);
          else c_qovf++;
          if (type != RAY_SHADOW && slot >= 0) queue_store(q.fin, (unsigned)slot, r, slot);
          active = false;
        } else {
          float t;
          const int next = tlas_route(sd.tlas, o, d, t_max, inst, false, t);
          if (next < 0) {
            if (type == RAY_SHADOW) { // TracerBase.h:396-399
              Ray r;
              reload(r);
              if (xlength(r.color) > 0.f) {
                fb_add(sd.fb, r.id, xmul(r.color, r.w));
                c_fb++;
              }
            }
            active = false;
          } else {
            o = xadd(o, xmul(d, xmul(t, 0.95f))); // TracerBase.h:394
            if (to_light && type == RAY_SHADOW) t_max = xsub(t_max, xmul(t, 0.95f)); // the distance left to the light
            const int owner = fp.world > 1 ? sd.inst[next].owner : fp.rank;
            if (owner != fp.rank) {
              Ray r;
              reload(r);
              if (q.peer_inbox) { // SendRays over NVLink: slot reservation + SoA stores in the owner's memory
                const RayQueue dst = q.peer_inbox[owner];
                const unsigned pos = queue_reserve_peer(dst.count, owner);
                if (pos < dst.cap)
                  queue_store(dst, pos, r, next);
                else
                  c_qovf++;
              } else {
                const unsigned pos = queue_reserve(q.exq.count);
                if (pos < q.exq.cap) queue_store(q.exq, pos, r, next);
                else c_qovf++;
              }
              c_exp++;
              active = false;
            } else {
              inst = next;
              begin_instance();
            }
          }
        }
      }
    }
  }
  if (stack.overflow) c_ovf = 1;
  count_add(&sd.ctr->instance_visits, c_visits);
  count_add(&sd.ctr->shadow_unoccluded, c_shok);
  count_add(&sd.ctr->fb_adds, c_fb);
  count_add(&sd.ctr->rays_exported, c_exp);
  count_add(&sd.ctr->stack_overflows, c_ovf);
  count_add(&sd.ctr->queue_overflows, c_qovf);
  if (gen) {
    count_add(&sd.ctr->routed_initial, c_routed);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&sd.ctr->camera_rays, (unsigned long long)n);
  }
  if (COUNT) {
    count_add(&sd.ctr->nodes_visited, tc.nodes);
    count_add(&sd.ctr->tris_tested, tc.tris);
  }
}

// (4 CTAs per SM = 64 registers: sibenik's two k_shade launches 0.22 -> 0.19 ms, but the incoherent 4K / 9 spp frame, where k_shade
// streams 50 M hits per wave from HBM, 23.7 -> 26.2 ms: 3 it is)
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 3
#endif
// k_shade: one thread per closest hit.  Surface data, Shade() per light -> shadow rays into `shq`, the bounce decision ->
// the secondary ray into `nxq` (EmbreeMeshAdapter.cpp:491-601).  Fully coherent: every lane of a warp shades.
__global__ void __launch_bounds__(256, SHADE_MIN_BLOCKS) k_shade(FrameParams fp, SceneDev sd, ShadeQueues q, int mode) {
  const unsigned n = min(*q.hq.count, q.hq.cap);
  const bool to_light = (fp.flags & FLAG_SHADOW_TO_LIGHT) != 0;
  unsigned long long c_hits = 0, c_shgen = 0, c_sec = 0, c_qovf = 0;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Ray r;
    int inst;
    queue_load(q.hq, i, r, inst);
    const int4 hc = q.hit[i];
    Hit h;
    h.prim = hc.x, h.slot = hc.y, h.t = r.t, h.u = u2f((uint32_t)hc.z), h.v = u2f((uint32_t)hc.w);
    const int slot = inst; // adapter mode only
    if (mode == 1) inst = 0;
    const Instance in = sd.inst[inst];
    const Blas &bl = sd.blas[in.mesh];
    c_hits++;
    V3 normal;
    Mat mat;
    surface_at(bl, in, h, r, fp.flags, normal, mat);
    if (r.type == RAY_SECONDARY) { // :572-575
      float tt = (h.t > 1.f) ? xdiv(1.f, h.t) : h.t;
      r.w = xmul(r.w, tt);
    }
    for (int li = 0; li < fp.n_lights; ++li) {
      const Light &L = fp.lights[li];
      Ray sh;
      if (!make_shadow_ray(r, normal, mat, L, light_position(L, fp.seed, r.id, r.depth, li), to_light, sh)) continue;
      c_shgen++;
      unsigned pos = queue_reserve(q.shq.count);
      if (pos < q.shq.cap) queue_store(q.shq, pos, sh, mode == 1 ? -1 : inst);
      else c_qovf++;
    }
    if (make_bounce(r, normal, fp.seed)) {
      c_sec++;
      unsigned pos = queue_reserve(q.nxq.count);
      if (pos < q.nxq.cap) queue_store(q.nxq, pos, r, mode == 1 ? slot : inst);
      else c_qovf++;
    } else if (mode == 1 && slot >= 0) {
      queue_store(q.fin, (unsigned)slot, r, slot);
    }
  }
  count_add(&sd.ctr->primary_hits, c_hits);
  count_add(&sd.ctr->shadow_generated, c_shgen);
  count_add(&sd.ctr->secondary_generated, c_sec);
  count_add(&sd.ctr->queue_overflows, c_qovf);
}

// closest-hit / any-hit queries in mesh-local space (parity hooks)
__global__ void k_intersect(Blas bl, RayQueue in, unsigned n, int any_hit, int4 *hits) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ray r;
  int dummy;
  queue_load(in, i, r, dummy);
  Hit h;
  TravCounters tc;
  if (any_hit)
    blas_traverse<true, false>(bl, r.o, r.d, GVT_RAY_EPSILON, FLT_MAX, h, &tc);
  else
    blas_traverse<false, false>(bl, r.o, r.d, GVT_RAY_EPSILON, FLT_MAX, h, &tc);
  hits[i] = make_int4(h.prim, (int)f2u(h.prim >= 0 ? h.t : FLT_MAX), (int)f2u(h.u), (int)f2u(h.v));
}

// AoS (80-byte host records) <-> SoA columns
__global__ void k_aos_to_soa(const uint4 *aos, unsigned n, RayQueue q, unsigned offset, int inst) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 *p = aos + 5 * (size_t)i;
  uint4 a = p[0], b = p[1], c = p[2], d = p[3];
  q.c0[offset + i] = make_float4(u2f(a.x), u2f(a.y), u2f(a.z), u2f(a.w));
  q.c1[offset + i] = make_float4(u2f(b.x), u2f(b.y), u2f(b.z), u2f(b.w));
  q.c2[offset + i] = make_float4(u2f(c.x), u2f(c.y), u2f(c.z), u2f(c.w));
  q.c3[offset + i] = make_int4((int)d.x, (int)d.y, (int)d.z, (int)d.w);
  q.inst[offset + i] = (inst == -2) ? (int)i : inst; // -2: the column carries the caller's slot index (adapter mode)
}
__global__ void k_soa_to_aos(RayQueue q, unsigned n, uint4 *aos) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = q.c0[i], b = q.c1[i], c = q.c2[i];
  int4 d = q.c3[i];
  uint4 *p = aos + 5 * (size_t)i;
  p[0] = make_uint4(f2u(a.x), f2u(a.y), f2u(a.z), f2u(a.w));
  p[1] = make_uint4(f2u(b.x), f2u(b.y), f2u(b.z), f2u(b.w));
  p[2] = make_uint4(f2u(c.x), f2u(c.y), f2u(c.z), f2u(c.w));
  p[3] = make_uint4((unsigned)d.x, (unsigned)d.y, (unsigned)d.z, (unsigned)d.w);
  p[4] = make_uint4(0u, 0u, 0u, 0u);
}

// shuffleRays on a host-provided batch (parity path): route, advance, compact survivors; finished shadow rays -> fb
__global__ void k_shuffle(FrameParams fp, SceneDev sd, RayQueue in, unsigned n, int from, RayQueue out) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ray r;
  int dummy;
  queue_load(in, i, r, dummy);
  float t;
  int next = tlas_route(sd.tlas, r.o, r.d, r.t_max, from, from == -1 && (fp.flags & FLAG_ENTER_INSIDE) != 0, t);
  if (next >= 0) {
    r.o = xadd(r.o, xmul(r.d, xmul(t, 0.95f)));
    if ((fp.flags & FLAG_SHADOW_TO_LIGHT) && r.type == RAY_SHADOW) r.t_max = xsub(r.t_max, xmul(t, 0.95f));
    unsigned pos = queue_reserve(out.count);
    if (pos < out.cap) queue_store(out, pos, r, next);
  } else if (r.type == RAY_SHADOW && xlength(r.color) > 0.f) {
    fb_add(sd.fb, r.id, xmul(r.color, r.w));
  }
}

// ---- DomainScheduler exchange helpers (replace SendRays' pack/unpack, DomainTracer.h:370-496) ---------------
// Wire record = the reference's 80-byte Ray (64-byte payload + 16 pad bytes, DomainTracer.h:444-451); the first pad
// word carries the target instance id (the reference sends it once per queue as `int queue_id`).
__global__ void k_export_count(RayQueue ex, const Instance *inst, unsigned *peer_counts) {
  unsigned n = min(*ex.count, ex.cap);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    atomicAdd(&peer_counts[inst[ex.inst[i]].owner], 1u);
}
// pack peer-major; peer_cursor starts at the exclusive offsets of the peer counts
__global__ void k_export_pack(RayQueue ex, const Instance *inst, unsigned *peer_cursor, uint4 *packed) {
  unsigned n = min(*ex.count, ex.cap);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int target = ex.inst[i];
    unsigned pos = atomicAdd(&peer_cursor[inst[target].owner], 1u);
    uint4 *p = packed + 5 * (size_t)pos;
    p[0] = reinterpret_cast<const uint4 *>(ex.c0)[i];
    p[1] = reinterpret_cast<const uint4 *>(ex.c1)[i];
    p[2] = reinterpret_cast<const uint4 *>(ex.c2)[i];
    p[3] = reinterpret_cast<const uint4 *>(ex.c3)[i];
    p[4] = make_uint4((unsigned)target, 0u, 0u, 0u);
  }
}
// unpack n wire records into the queue at [offset, offset+n)
__global__ void k_import(const uint4 *packed, unsigned n, RayQueue q, unsigned offset) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 *p = packed + 5 * (size_t)i;
  reinterpret_cast<uint4 *>(q.c0)[offset + i] = p[0];
  reinterpret_cast<uint4 *>(q.c1)[offset + i] = p[1];
  reinterpret_cast<uint4 *>(q.c2)[offset + i] = p[2];
  reinterpret_cast<uint4 *>(q.c3)[offset + i] = p[3];
  q.inst[offset + i] = (int)p[4].x;
}

// ---- NVLink peer path: flags, framebuffer gather / accumulate (replace icetCompositeImage, IceTComposite.cpp:96) ----------
// Every rank owns one Mailbox that its peers can write: `go` is set by the root when its framebuffer may be overwritten,
// done[r] by rank r when its contribution to this rank's framebuffer is complete.  Values are frame sequence numbers.
#define GVT_P2P_MAX 16
// Control block of the persistent DomainScheduler kernel (k_persist), one per rank, mapped into every peer.  The three hot
// words sit on separate 128-byte lines (remote atomics on `tail` must not fight the local atomics on `head`).
struct PersistCtl {
  unsigned tail; // inbox slots reserved so far by producers on any rank (== rays ever put into this inbox this frame)
  unsigned pad0[31];
  unsigned head;     // inbox tickets handed to consumer lanes of this rank
  unsigned pad1[31];
  unsigned consumed; // rays and hits this rank has retired
  unsigned err;      // 1 = a wait timed out, 2 = a queue overflowed (the frame is invalid either way)
  unsigned pad2[30];
  unsigned hit_tail; // closest hits queued for the shader SMs (reserved slots)
  unsigned pad3[31];
  unsigned hit_head; // hit-queue tickets handed out
  unsigned cam_warps; // warps that have finished phase 1 (camera rays)
  unsigned pad4[30];
  // counted frames only (k_persist<true>): SM cycles summed over warps -- [0] camera phase, tracer warps: [1] waiting for
  // rays, [2] in the traversal loop, [3] finishing rays (routing, sends, hit records), [4] drawing / polling tickets;
  // shader warps: [5] shading, [6] waiting for hits; [7] warps counted
  unsigned long long prof[16];
  // the root's control warp samples the job's progress while it polls for termination: [k] = {microseconds since the start
  // barrier, rays retired on all ranks, rays handed out on all ranks}, one sample every GVT_PROGRESS_US
  unsigned progress[3 * 256];
  unsigned n_progress, pad5[31];
  // ---- not reset between frames: values are frame sequence numbers ----
  unsigned cam_done; // every camera ray of this rank's shard has been handed out
  unsigned go, stop, done_seq;
  unsigned long long t_go, t_last; // globaltimer at the start barrier / at the last retire flush (per-rank busy time)
  unsigned arrive[GVT_P2P_MAX];    // start barrier: arrive[r] = frame number, written by rank r
};
#define GVT_PERSIST_RESET_BYTES (768 + 3 * 256 * 4 + 128) /* tail .. pad5 */
#define GVT_PROGRESS_US 16
struct Mailbox {
  unsigned go, err, cta_done, pad;
  unsigned done[GVT_P2P_MAX];
  unsigned reserved[60 - GVT_P2P_MAX];
  PersistCtl pc;
};
struct PeerMail {
  Mailbox *m[GVT_P2P_MAX];
};
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// thread i < n (i != skip) waits until flags[i] has reached sequence number seq; gives up after timeout_ns (err is set)
__global__ void k_p2p_wait(const unsigned *flags, int n, int skip, unsigned seq, unsigned *err, unsigned long long timeout_ns) {
  const int i = (int)threadIdx.x;
  if (i >= n || i == skip) return;
  const unsigned long long t0 = global_ns();
  while ((int)(ld_acquire_sys(flags + i) - seq) < 0) {
    __nanosleep(128);
    if (global_ns() - t0 > timeout_ns) {
      atomicExch(err, 1u);
      return;
    }
  }
}
// the root releases its peers: thread r stores seq into rank r's `go`
__global__ void k_p2p_go(PeerMail pm, int world, int self, unsigned seq) {
  const int r = (int)threadIdx.x;
  if (r < world && r != self) st_release_sys(&pm.m[r]->go, seq);
}
// last CTA of a kernel that wrote into the root's memory: publish done[rank] = seq at the root
__device__ __forceinline__ void p2p_signal_done(Mailbox *mine, Mailbox *root_mail, int rank, unsigned seq) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned d = atomicAdd(&mine->cta_done, 1u);
    if (d == gridDim.x - 1) {
      mine->cta_done = 0;
      __threadfence_system();
      st_release_sys(&root_mail->done[rank], seq);
    }
  }
}
// ImageScheduler: clamp this shard's film rows and store them into dst (the root's framebuffer; dst == src on the root)
__global__ void k_rows_to_root(const float4 *src, float4 *dst, RowTiles tiles, int W, Mailbox *mine, Mailbox *root_mail, int rank, unsigned seq) {
  const size_t n = tiles_local_rows(tiles) * (size_t)W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t lr = i / W, x = i - lr * W;
    const size_t b = lr / tiles.tile_rows, k = lr - b * tiles.tile_rows;
    const size_t pix = ((size_t)tiles.row0 + (b * tiles.stride + tiles.index) * tiles.tile_rows + k) * W + x;
    float4 v = src[pix];
    v.x = fminf(v.x, 1.f), v.y = fminf(v.y, 1.f), v.z = fminf(v.z, 1.f);
    dst[pix] = v;
  }
  if (root_mail) p2p_signal_done(mine, root_mail, rank, seq);
}
// this shard's film rows of `src` (a framebuffer in this GPU's memory or a peer's, mapped over NVLink), clamped, packed row
// tile after row tile into `dst`: the staging buffer of the parallel read-back (every rank brings its own rows to the host
// over its own PCIe link; replaces the gather that follows icetCompositeImage, composite/IceTComposite.cpp:96-101)
__global__ void k_rows_pack(const float4 *src, float4 *dst, RowTiles tiles, int W) {
  const size_t n = tiles_local_rows(tiles) * (size_t)W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t lr = i / W, x = i - lr * W;
    const size_t b = lr / tiles.tile_rows, k = lr - b * tiles.tile_rows;
    const size_t pix = ((size_t)tiles.row0 + (b * tiles.stride + tiles.index) * tiles.tile_rows + k) * W + x;
    float4 v = __ldcg(src + pix);
    v.x = fminf(v.x, 1.f), v.y = fminf(v.y, 1.f), v.z = fminf(v.z, 1.f);
    dst[i] = v;
  }
}
// DomainScheduler: add the pixels this rank touched (alpha counts contributions) to the root's framebuffer
__global__ void k_fb_accumulate_root(const float4 *src, float *dst, size_t npix, Mailbox *mine, Mailbox *root_mail, int rank, unsigned seq) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    if (v.w != 0.f) {
      atomicAdd_system(dst + 4 * i, v.x), atomicAdd_system(dst + 4 * i + 1, v.y);
      atomicAdd_system(dst + 4 * i + 2, v.z), atomicAdd_system(dst + 4 * i + 3, v.w);
    }
  }
  p2p_signal_done(mine, root_mail, rank, seq);
}


// =========================================================================================================
// The persistent DomainScheduler kernel.  Tracer<DomainScheduler>::operator() (algorithm/DomainTracer.h:185-368) + SendRays
// (:370-496) + the termination test (:339-349) + the composite (composite/IceTComposite.cpp:84-117) as ONE kernel launch per
// rank and frame; the reference's asynchronous loop (tracer/Domain/DomainTracer.cpp:109-192, scheduler::AsyncDomain) is the
// closest relative: there are no exchange rounds.
//
//   phase 1      every CTA: this rank's 1/world of the camera rays (row tiles dealt round-robin) are generated, routed through
//                the instance BVH and stored into the INBOX of a rank that serves their first instance.
//   phase 2      SMs split into two roles (by %smid, so that an SM's instruction cache holds one role's code: the first
//                version ran everything in every warp, 133 KB of SASS, and stalled on instruction fetch three times per issue):
//     tracer SMs   lanes draw tickets from the rank's inbox and wait for their slot; trace in the instance (closest hit / any
//                  hit); miss -> next instance through the TLAS: traced here if this rank serves it, else stored into a
//                  serving rank's inbox; closest hit -> hit record into the rank's HIT QUEUE; a finished shadow ray adds its
//                  colour to the ROOT's framebuffer (remote atomics over NVLink).
//     shader SMs   lanes draw tickets from the hit queue: surface, Shade() per light, shadow rays and the bounce go into the
//                  rank's own inbox (EmbreeMeshAdapter.cpp:491-601).
//   queues       an inbox is a slot array in its owner's HBM that every peer has mapped (CUDA IPC / peer access).  A producer
//                on any GPU reserves a slot with one atomic on `tail`, stores the SoA ray and publishes it by writing the frame
//                number into the slot's `ready` word (release); consumers draw tickets from `head` and poll their slot's
//                `ready` word.  No kernel launch, host round trip or collective sits between a ray leaving one GPU and being
//                traced on the next.
//   termination  every rank counts what it retires (`consumed`); `tail` + `hit_tail` is what it was given.  One control warp
//                on the root reads all `consumed` words, then all `given` words: equal sums (once every rank has finished phase
//                1) mean that at the instant between the two passes nothing was queued or in flight anywhere (both counts only
//                grow, and a ray is counted as given before its parent is counted as retired).  The root then releases every
//                rank and clamps the image.
//   servers      an instance may be served by several ranks (its owner plus ranks that hold a replica of its BLAS or have
//                the owner's BLAS mapped over NVLink); a ray picks one by a hash of its direction bits against the
//                cumulative weights of the table (gvtb200_domain_set_servers) -- the load balance the reference's hybrid
//                schedulers looked for by loading hot domains on idle ranks.
// Every wait gives up after timeout_ns and flags the frame as failed.
// =========================================================================================================
#ifndef PERSIST_MIN_BLOCKS
#define PERSIST_MIN_BLOCKS 6
#endif
struct PersistParams {
  RayQueue inbox;             // this rank's inbound slots (count unused: PersistCtl::tail / head)
  RayQueue hitq;              // this rank's closest hits: ray + hit column, ticketed like the inbox (hit_tail / hit_head)
  int4 *hit;                  // hit column of hitq: prim, triangle slot, u bits, v bits
  const RayQueue *peer_inbox; // [world]: every rank's inbox as mapped into this GPU (entry `rank` == inbox)
  PeerMail mail;              // every rank's mailbox (PersistCtl inside)
  float *root_fb;             // the root's RGBA framebuffer as mapped here
  const uint4 *servers;       // per instance {ranks (4 x 8 bit), thresholds 0/1 (2 x 16 bit), threshold 2, count} or null
  unsigned long long *inst_load; // COUNT only: wide nodes visited per instance (input of the replica planner)
  unsigned long long timeout_ns;
  RowTiles tiles;     // this rank's shard of the film rows
  unsigned cam_count; // camera rays of the shard
  unsigned seq;       // frame number (>= 1, the same on every rank)
  int rank, world, root;
  int shade_sms;      // SMs with %smid below this shade, the others trace; 0: every shade_mod-th CTA shades (bounded test grids)
  int shade_mod;
  int clamp;          // the root clamps the image when the frame is over
  int timeline;       // record when this rank last retired a ray (per-rank busy time)
  size_t fb_pix;
};

// Polling loads and publishing stores of the worker warps.  ld.acquire / __threadfence() compile to LDG.STRONG + CCTL.IVALL:
// they would wipe the SM's L1 (the BVH nodes every other warp is traversing) at every poll.  A worker therefore polls with
// ld.volatile (LDG.STRONG.SYS, no invalidation), reads what the flag guards with ld.cg (L2, the coherence point for local
// and peer writers alike; issued after the flag value has come back, the SM does not speculate loads), and publishes with
// st.release / red.release (MEMBAR.ALL + the store: orders the lane's earlier writes, no invalidation).
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// What a `ready` word guards is read with volatile asm loads as well: a plain __ldcg may be hoisted by the compiler above the
// test of the flag value (it was, once the flag was polled ahead of the traversal burst: the payload was then fetched
// together with the flag and a slot that had just been reserved was read before it was written).  asm volatile keeps the
// loads behind the branch, LDG.STRONG reads L2.
__device__ __forceinline__ float4 ld_strong_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int4 ld_strong_i4(const int4 *p) {
  int4 v;
  asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_strong_i32(const int *p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// release as ONE fence and any number of relaxed stores (st.release is MEMBAR + store, once per flag)
__device__ __forceinline__ void fence_release(bool far) {
  if (far) asm volatile("fence.acq_rel.sys;" ::: "memory"); else asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned *p, unsigned v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_gpu(unsigned *p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release_sys_add(unsigned *p, unsigned v) { asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned sm_id() {
  unsigned v;
  asm("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
__device__ __forceinline__ int pick_server(const uint4 *servers, int inst, int owner, V3 d) {
  if (!servers) return owner;
  const uint4 e = __ldg(servers + inst);
  if (e.w <= 1u) return (int)(e.x & 0xFFu);
  const unsigned h = pcg(f2u(d.x) + 0x9e3779b9u * f2u(d.y) + 0x85ebca6bu * f2u(d.z)) >> 16;
  const unsigned j = (h >= (e.y & 0xFFFFu) ? 1u : 0u) + (h >= (e.y >> 16) ? 1u : 0u) + (h >= e.z ? 1u : 0u);
  return (int)((e.x >> (8u * min(j, e.w - 1u))) & 0xFFu);
}

// the control warp (block 0, warp 0): start barrier, termination, release
__device__ __noinline__ void persist_control(const PersistParams &pp, PersistCtl *ctl) {
  const unsigned FULL = 0xffffffffu;
  const int r = (int)threadIdx.x;
  const bool peer = r < pp.world;
  const unsigned seq = pp.seq;
  const unsigned long long t0 = global_ns();
  PersistCtl *theirs = peer ? &pp.mail.m[r]->pc : ctl;
  // start barrier: every rank has reset its counters (and the root has cleared the image) before anybody produces
  bool bad = false;
  if (peer) {
    st_release_sys(&theirs->arrive[pp.rank], seq);
    while ((int)(ld_volatile_u32(&ctl->arrive[r]) - seq) < 0) {
      __nanosleep(64);
      if (global_ns() - t0 > pp.timeout_ns) {
        bad = true;
        break;
      }
    }
  }
  bad = __any_sync(FULL, bad);
  if (r == 0) {
    ctl->t_go = global_ns();
    if (bad) atomicOr(&ctl->err, 1u), st_release_sys(&ctl->stop, seq);
    st_release_sys(&ctl->go, seq);
  }
  if (bad) return;
  if (pp.rank == pp.root) {
    for (;;) {
      // pass 1: what every rank has retired (and whether its camera rays are all out); pass 2 (issued after every pass-1
      // value has come back: the reduction below consumes them): what every rank was given.  Volatile loads read the
      // owner's L2, where its atomics are performed.
      const unsigned c = peer ? ld_volatile_u32(&theirs->consumed) : 0u;
      const unsigned e = peer ? ld_volatile_u32(&theirs->err) : 0u;
      const unsigned started = peer ? (ld_volatile_u32(&theirs->cam_done) == seq ? 1u : 0u) : 1u;
      const unsigned cs = __reduce_add_sync(FULL, c), es = __reduce_or_sync(FULL, e), all_started = __reduce_and_sync(FULL, started);
      const unsigned t = (peer && cs != 0xFFFFFFFFu) ? ld_volatile_u32(&theirs->tail) + ld_volatile_u32(&theirs->hit_tail) : 0u;
      const unsigned ts = __reduce_add_sync(FULL, t);
      const unsigned long long now = global_ns();
      const bool late = now - t0 > pp.timeout_ns;
      if (r == 0) { // progress curve of the whole job (read back by gvtb200_domain_persist_progress)
        const unsigned us = (unsigned)((now - ctl->t_go) / 1000ull), k = ctl->n_progress;
        if (k < 256u && us >= k * GVT_PROGRESS_US) ctl->progress[3 * k] = us, ctl->progress[3 * k + 1] = cs, ctl->progress[3 * k + 2] = ts, ctl->n_progress = k + 1u;
      }
      if ((all_started && cs == ts) || es || late) {
        if ((es || late) && r == 0) atomicOr(&ctl->err, es ? es : 1u);
        break;
      }
      __nanosleep(400);
    }
    if (peer) st_release_sys(&theirs->done_seq, seq);
  }
  if (r == 0) {
    while ((int)(ld_volatile_u32(&ctl->done_seq) - seq) < 0) {
      __nanosleep(100);
      if (global_ns() - t0 > pp.timeout_ns) {
        atomicOr(&ctl->err, 1u);
        break;
      }
    }
    st_release_sys(&ctl->stop, seq);
  }
}

template <bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS, PERSIST_MIN_BLOCKS) k_persist(FrameParams fp, SceneDev sd, PersistParams pp, TraceTune tune) {
  __shared__ uint2 s_stack[TRACE_SM_STACK * TRACE_THREADS];
  __shared__ unsigned char s_lut[8 * 256];
  for (unsigned k = threadIdx.x; k < 8u * 256u; k += TRACE_THREADS) s_lut[k] = (unsigned char)best_slot(k & 255u, k >> 8);
  __syncthreads(); // the only block-wide barrier: the control warp leaves the block's flow after it
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31, ltmask = (1u << lane) - 1u;
  PersistCtl *ctl = &pp.mail.m[pp.rank]->pc;
  const unsigned seq = pp.seq;
  const int me = pp.rank;
  const bool multi = pp.world > 1;
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    persist_control(pp, ctl);
    return;
  }
  { // wait until every rank is ready
    const unsigned long long t0 = global_ns();
    if (lane == 0)
      while (ld_volatile_u32(&ctl->go) != seq) {
        __nanosleep(64);
        if (global_ns() - t0 > 2 * pp.timeout_ns) break;
      }
    __syncwarp();
  }
  const bool to_light = (fp.flags & FLAG_SHADOW_TO_LIGHT) != 0;
  unsigned long long c_visits = 0, c_shok = 0, c_fb = 0, c_exp = 0, c_ovf = 0, c_qovf = 0, c_routed = 0, c_hits = 0, c_shgen = 0, c_sec = 0;
  unsigned c_cons = 0; // retired since the last flush
  bool dirty = false;  // this lane has stored into another GPU's memory since the last flush

  // A stored ray becomes visible when its slot's `ready` word is written with release semantics (MEMBAR.ALL + store): done
  // right after the payload stores, the barrier waits for their acknowledgement -- an NVLink round trip for a remote inbox,
  // during which the other 31 lanes of the warp stand still.  The word is therefore written one or two loop iterations
  // later, when the payload has long arrived and the barrier finds nothing outstanding.
  unsigned *pend0 = nullptr, *pend1 = nullptr; // (scalars: a dynamically indexed array would live in local memory)
  bool far0 = false, far1 = false;
  auto publish = [&]() {
    if (pend0) {
      fence_release(far0 || (pend1 && far1)); // one fence orders the payload of both
      if (far0) st_relaxed_sys(pend0, seq); else st_relaxed_gpu(pend0, seq);
      pend0 = nullptr;
      if (pend1) {
        if (far1) st_relaxed_sys(pend1, seq); else st_relaxed_gpu(pend1, seq);
        pend1 = nullptr;
      }
    }
  };
  auto pend_add = [&](unsigned *flag, bool far) {
    if (pend0 && pend1) publish();
    if (!pend0) pend0 = flag, far0 = far; else pend1 = flag, far1 = far;
  };
  // SendRays: into the inbox of `dest` (this rank's own inbox for the rays it makes)
  auto send = [&](const Ray &r, int target, int dest) {
    const RayQueue q = dest == me ? pp.inbox : pp.peer_inbox[dest];
    PersistCtl *pc = &pp.mail.m[dest]->pc;
    unsigned pos;
    if (dest == me) {
      pos = queue_reserve(&pc->tail);
    } else {
      pos = queue_reserve_peer(&pc->tail, dest);
      dirty = true;
    }
    if (pos < q.cap) {
      queue_store(q, pos, r, target);
      pend_add(q.ready + pos, dest != me);
    } else { // never silently: the frame fails with GVTB200_ERR_CAPACITY; the count stays balanced so that it still ends
      c_qovf++;
      atomicOr(&ctl->err, 2u);
      if (dest == me) atomicAdd(&pc->consumed, 1u); else atomicAdd_system(&pc->consumed, 1u);
    }
  };
  // retired rays -> `consumed`; everything this warp wrote anywhere (peer inboxes, the root's image) is ordered before it.
  // One release-reduction per warp (same-address atomics serialise in L2 at ~1 per clock: one per lane and retired ray cost
  // more than the traversal of a mostly-empty frame).  The warp barrier orders the other lanes' writes before lane 0's
  // release (causality order is cumulative in the PTX memory model).
  auto flush = [&]() {
    const unsigned tot = __reduce_add_sync(FULL, c_cons);
    const bool far = __any_sync(FULL, dirty);
    c_cons = 0u, dirty = false;
    if (tot != 0u && lane == 0) {
      if (far) red_release_sys_add(&ctl->consumed, tot); else red_release_gpu_add(&ctl->consumed, tot);
      if (pp.timeline) atomicMax(&ctl->t_last, global_ns());
    }
  };

  // ---- phase 1: gvtPerspectiveCamera::generateRays + FilterRaysLocally (first routing hop) for this rank's film shard ------
  // Two passes over chunks of CAM_CHUNK rays per lane: (A) generate, route, pick the serving rank, count per destination
  // inside the CTA; ONE slot reservation per CTA and destination (same-address atomics on an inbox's `tail` serialise at
  // a few nanoseconds each over NVLink, and one per ray group made this phase 0.3 ms long whatever the number of GPUs); (B) regenerate the ray (cheaper than keeping it) and store it into its slot.
  long long pc0 = COUNT ? clock64() : 0;
  {
    constexpr int CAM_CHUNK = 4;
    __shared__ unsigned s_base[GVT_P2P_MAX]; // per CTA and destination: where this chunk's slots start
    // rays per (pass j, warp, destination), then their exclusive prefix in that order: a destination's slots are filled in FILM
    // ORDER, so that the 32 tickets a tracer warp draws are 32 neighbouring camera rays, as they are for k_trace (slots handed
    // out by an atomic in arrival order mixed four distant film regions in every warp: one GPU, 4 rays per pixel 24.8 ms
    // against 18.0 ms for the wave path)
    __shared__ unsigned s_wcnt[CAM_CHUNK * (TRACE_THREADS / 32) * GVT_P2P_MAX];
    const unsigned warp = threadIdx.x >> 5;
    const unsigned n_workers = blockIdx.x == 0 ? TRACE_THREADS - 32 : TRACE_THREADS, first_worker = blockIdx.x == 0 ? 32u : 0u;
    auto cta_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(n_workers) : "memory"); }; // the control warp is not part of it
    const bool enter_inside = (fp.flags & FLAG_ENTER_INSIDE) != 0;
    const size_t per_row = (size_t)fp.cam.W * fp.cam.samples * fp.cam.samples;
    const size_t n_warps = (size_t)gridDim.x * (TRACE_THREADS / 32) - 1, wg = (size_t)blockIdx.x * (TRACE_THREADS / 32) + warp - 1; // workers only
    auto cam_index = [&](size_t i) -> size_t { // shard ray i -> global camera-ray index (RowTiles arithmetic)
      const size_t lr = i / per_row, within = i - lr * per_row;
      const size_t bb = lr / pp.tiles.tile_rows, k = lr - bb * pp.tiles.tile_rows;
      return ((size_t)pp.tiles.row0 + (bb * pp.tiles.stride + pp.tiles.index) * pp.tiles.tile_rows + k) * per_row + within;
    };
    // per-lane notes of a chunk live in the traversal stack's shared memory (unused until phase 2), three 8-byte entries per
    // ray: {first instance, destination | position in the warp's reservation << 8}, {d.x, d.y}, {d.z, routing distance}
    uint2 *note = s_stack + threadIdx.x;
    static_assert(3 * CAM_CHUNK <= TRACE_SM_STACK, "camera notes must fit the shared stack");
    // A camera ray is a function of its index and the camera every rank holds: another rank gets the index, not the record --
    // the first hop of 7/8 of the primary rays is what fills the NVLink ports at the start of a frame
    const bool cam_ref = pp.fb_pix * (size_t)(fp.cam.samples * fp.cam.samples) < 0x7FFFFFFFull;
    for (size_t it0 = 0; (it0 * n_warps) * 32 < pp.cam_count; it0 += CAM_CHUNK) {
      for (unsigned k = threadIdx.x - first_worker; k < (unsigned)(CAM_CHUNK * (TRACE_THREADS / 32) * GVT_P2P_MAX); k += n_workers) s_wcnt[k] = 0u;
      cta_sync();
#pragma unroll 1
      for (int j = 0; j < CAM_CHUNK; ++j) {
        const size_t i = ((it0 + j) * n_warps + wg) * 32 + lane;
        uint2 nt = make_uint2(0xFFFFFFFFu, 0u); // background / beyond the shard
        unsigned to = 0xFFu;                    // destination rank, 0xFF = none
        if (i < pp.cam_count) {
          Ray r;
          camera_ray(fp.cam, cam_index(i), r);
          float t;
          const int next = tlas_route(sd.tlas, r.o, r.d, r.t_max, -1, enter_inside, t);
          if (next >= 0) {
            c_routed++;
            const int dest = multi ? pick_server(pp.servers, next, sd.inst[next].owner, r.d) : me;
            if (dest != me) c_exp++;
            nt.x = (unsigned)next, to = (unsigned)dest;
            note[(3 * j + 1) * TRACE_THREADS] = make_uint2(f2u(r.d.x), f2u(r.d.y));
            note[(3 * j + 2) * TRACE_THREADS] = make_uint2(f2u(r.d.z), f2u(t));
          }
        }
        const unsigned same = __match_any_sync(FULL, to); // lanes of this warp bound for the same rank, in lane order
        if (to != 0xFFu) {
          nt.y = to | ((unsigned)__popc(same & ltmask) << 8);
          if ((same & ltmask) == 0u) s_wcnt[((unsigned)j * (TRACE_THREADS / 32) + warp) * GVT_P2P_MAX + to] = (unsigned)__popc(same);
        }
        note[3 * j * TRACE_THREADS] = nt;
      }
      cta_sync();
      if ((int)(threadIdx.x - first_worker) < pp.world) { // ONE reservation per CTA, chunk and destination
        const int dst = (int)(threadIdx.x - first_worker);
        unsigned n = 0;
        for (unsigned k = 0; k < (unsigned)(CAM_CHUNK * (TRACE_THREADS / 32)); ++k) {
          const unsigned c = s_wcnt[k * GVT_P2P_MAX + dst];
          s_wcnt[k * GVT_P2P_MAX + dst] = n;
          n += c;
        }
        if (n) s_base[dst] = dst == me ? atomicAdd(&ctl->tail, n) : atomicAdd_system(&pp.mail.m[dst]->pc.tail, n);
      }
      cta_sync();
#pragma unroll 1
      for (int j = 0; j < CAM_CHUNK; ++j) {
        const uint2 nt = note[3 * j * TRACE_THREADS];
        if (nt.x == 0xFFFFFFFFu) continue;
        const size_t i = ((it0 + j) * n_warps + wg) * 32 + lane;
        const uint2 n1 = note[(3 * j + 1) * TRACE_THREADS], n2 = note[(3 * j + 2) * TRACE_THREADS];
        Ray r; // gvtPerspectiveCamera::generateRays' record (gvtCamera.cpp:279-300) around the direction pass A computed
        r.d = v3(u2f(n1.x), u2f(n1.y), u2f(n2.x));
        r.o = xadd(fp.cam.eye, xmul(r.d, xmul(u2f(n2.y), 0.95f))); // TracerBase.h:394
        r.t_min = GVT_RAY_EPSILON, r.t_max = FLT_MAX, r.color = v3(0.f, 0.f, 0.f), r.t = FLT_MAX, r.w = fp.cam.contri;
        r.type = RAY_PRIMARY, r.depth = fp.cam.depth, r.id = (int)(cam_index(i) / (size_t)(fp.cam.samples * fp.cam.samples));
        const int dest = (int)(nt.y & 0xFFu);
        const RayQueue q = dest == me ? pp.inbox : pp.peer_inbox[dest];
        const unsigned pos = s_base[dest] + s_wcnt[((unsigned)j * (TRACE_THREADS / 32) + warp) * GVT_P2P_MAX + dest] + (nt.y >> 8);
        note[3 * j * TRACE_THREADS] = make_uint2(nt.x, (unsigned)dest | ((pos - s_base[dest]) << 8)); // (the flag pass below)
        if (pos < q.cap) {
          if (dest == me || !cam_ref) {
            queue_store(q, pos, r, (int)nt.x);
          } else { // 20 bytes over NVLink instead of 68: the index, the routing distance and the instance (see start_landed)
            q.c3[pos] = make_int4((int)(unsigned)cam_index(i), (int)n2.y, 0, RAY_CAMERA_REF);
            q.inst[pos] = (int)nt.x;
          }
          if (dest != me) dirty = true;
        } else { // the frame fails with GVTB200_ERR_CAPACITY; the count stays balanced so that it still ends
          c_qovf++;
          atomicOr(&ctl->err, 2u);
          if (dest == me) atomicAdd(&ctl->consumed, 1u); else atomicAdd_system(&pp.mail.m[dest]->pc.consumed, 1u);
          note[3 * j * TRACE_THREADS] = make_uint2(0xFFFFFFFFu, 0u);
        }
      }
      __syncwarp();
      fence_release(dirty); // one fence per chunk waits for this lane's payload stores; the flags are relaxed stores after it
#pragma unroll 1
      for (int j = 0; j < CAM_CHUNK; ++j) {
        const uint2 nt = note[3 * j * TRACE_THREADS];
        if (nt.x == 0xFFFFFFFFu) continue;
        const int dest = (int)(nt.y & 0xFFu);
        if (dest == me) st_relaxed_gpu(pp.inbox.ready + s_base[dest] + (nt.y >> 8), seq);
        else st_relaxed_sys(pp.peer_inbox[dest].ready + s_base[dest] + (nt.y >> 8), seq);
      }
      cta_sync(); // s_cnt / s_base are rewritten by the next chunk
    }
    __syncwarp();
    // (the warp count below only tells the root that no more camera rays will be handed out by this rank)
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(&ctl->cam_warps, 1u) == (unsigned)n_warps - 1u) st_release_sys(&ctl->cam_done, seq);
    }
    __syncwarp();
  }

  if (COUNT && lane == 0) atomicAdd(&ctl->prof[0], (unsigned long long)(clock64() - pc0)), atomicAdd(&ctl->prof[7], 1ull);
  long long p_wait = 0, p_trav = 0, p_post = 0, p_poll = 0; // COUNT: cycles of this warp per activity
  const bool shader = pp.shade_sms > 0 ? (int)sm_id() < pp.shade_sms : (blockIdx.x % (unsigned)pp.shade_mod) == 1u;
  unsigned iter = 0;
  unsigned long long t_idle0 = 0;
  const unsigned long long t_start = global_ns();
  enum { L_IDLE = 0, L_WAIT = 1, L_ACTIVE = 2, L_DEAD = 3, L_HELP = 4, L_JOIN = 5, L_SEND = 6 };
  int st = L_IDLE;
  unsigned idx = 0; // the lane's ticket: inbox slot (tracer) or hit-queue slot (shader)

  if (shader) {
    // ================= shader SMs: EmbreeMeshAdapter.cpp:491-601 for every closest hit of this rank ==========================
    for (;;) {
      ++iter;
      if (COUNT) pc0 = clock64();
      const unsigned idle = __ballot_sync(FULL, st == L_IDLE);
      if (idle != 0u) { // tickets are drawn ahead of the hits, like the tracers' (one per lane at most)
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&ctl->hit_head, (unsigned)__popc(idle));
        base = __shfl_sync(FULL, base, 0);
        if (st == L_IDLE) idx = base + (unsigned)__popc(idle & ltmask), st = L_WAIT;
      }
      bool got = false;
      if (st == L_WAIT) {
        if (idx >= pp.hitq.cap) st = L_DEAD;
        else got = ld_volatile_u32(pp.hitq.ready + idx) == seq;
      }
      if (__ballot_sync(FULL, got) == 0u) {
        flush();
        unsigned stop = 0;
        if (lane == 0) stop = ld_volatile_u32(&ctl->stop) == seq ? 1u : 0u;
        if (__shfl_sync(FULL, stop, 0)) break;
        __nanosleep(256);
        if ((iter & 255u) == 0u) {
          const unsigned long long now = __shfl_sync(FULL, global_ns(), 0);
          if (t_idle0 == 0) t_idle0 = now;
          if (now - t_idle0 > 2 * pp.timeout_ns) {
            if (lane == 0) atomicOr(&ctl->err, 1u);
            break;
          }
        }
        if (COUNT) p_wait += clock64() - pc0;
        continue;
      }
      t_idle0 = 0;
      if (got) {
        Ray r;
        {
          const float4 a = ld_strong_f4(pp.hitq.c0 + idx), b = ld_strong_f4(pp.hitq.c1 + idx), c = ld_strong_f4(pp.hitq.c2 + idx);
          const int4 e = ld_strong_i4(pp.hitq.c3 + idx);
          r.o = v3(a.x, a.y, a.z), r.t_min = a.w, r.d = v3(b.x, b.y, b.z), r.t_max = b.w;
          r.color = v3(c.x, c.y, c.z), r.t = c.w;
          r.id = e.x, r.depth = e.y, r.w = u2f((uint32_t)e.z), r.type = e.w;
        }
        const int inst = ld_strong_i32(pp.hitq.inst + idx);
        const int4 hc = ld_strong_i4(pp.hit + idx);
        Hit h;
        h.prim = hc.x, h.slot = hc.y, h.t = r.t, h.u = u2f((uint32_t)hc.z), h.v = u2f((uint32_t)hc.w);
        const Instance in = sd.inst[inst];
        const Blas &bl = sd.blas[in.mesh];
        c_hits++;
        V3 normal;
        Mat mat;
        surface_at(bl, in, h, r, fp.flags, normal, mat);
        if (r.type == RAY_SECONDARY) { // :572-575
          const float tt = (h.t > 1.f) ? xdiv(1.f, h.t) : h.t;
          r.w = xmul(r.w, tt);
        }
        for (int li = 0; li < fp.n_lights; ++li) {
          const Light &L = fp.lights[li];
          Ray sh;
          if (!make_shadow_ray(r, normal, mat, L, light_position(L, fp.seed, r.id, r.depth, li), to_light, sh)) continue;
          c_shgen++;
          send(sh, inst, me); // the shadow ray starts in the instance that was hit
        }
        if (make_bounce(r, normal, fp.seed)) {
          c_sec++;
          send(r, inst, me);
        }
        c_cons++;
        st = L_IDLE;
      }
      publish();
      flush();
      if (COUNT) p_trav += clock64() - pc0;
    }
    publish();
    if (COUNT && lane == 0) atomicAdd(&ctl->prof[5], (unsigned long long)p_trav), atomicAdd(&ctl->prof[6], (unsigned long long)p_wait), atomicAdd(&ctl->prof[8], 1ull);
  } else {
    // ================= tracer SMs ========================================================================================
    TravCounters tc;
    tc.nodes = 0, tc.tris = 0;
    unsigned nodes0 = 0;
    SmemStack stack;
    stack.s = s_stack + threadIdx.x;
    stack.overflow = false;
    bool first = true, any = false;
    int inst = 0, type = 0;
    int team = -1; // helper lanes: the owner's lane | the lane's own state << 8
    V3 o = v3(0.f, 0.f, 0.f), d = o;
    float t_min = 0.f, t_max = 0.f, tfar = FLT_MAX;
    const float4 *nodes = nullptr, *tris = nullptr;
    RayLocal rl;
    Trav tv;
    ray_local(rl, o, o);
    trav_begin(tv, 0u, 0.f);
    auto begin_instance = [&]() {
      const Instance in = sd.inst[inst];
      const Blas &bl = sd.blas[in.mesh];
      V3 lo, ld;
      to_local(in, o, d, lo, ld);
      ray_local(rl, lo, ld);
      nodes = bl.nodes, tris = bl.tris;
      tfar = (any && to_light) ? t_max : FLT_MAX; // prepGVT_EMBREE_PACKET_TYPE (EmbreeMeshAdapter.cpp:269-286)
      trav_begin(tv, bl.n_tris, tfar);
      c_visits++;
      if (COUNT) nodes0 = tc.nodes;
    };
    auto reload = [&](Ray &r) { // the columns that are not kept in registers stay in the lane's inbox slot
      const float4 c2 = ld_strong_f4(pp.inbox.c2 + idx);
      const int4 c3 = ld_strong_i4(pp.inbox.c3 + idx);
      r.color = v3(c2.x, c2.y, c2.z), r.t = c2.w;
      r.id = c3.x, r.depth = c3.y, r.w = u2f((uint32_t)c3.z), r.type = c3.w;
      r.o = o, r.t_min = t_min, r.d = d, r.t_max = t_max;
    };
    unsigned landed = 0u; // L_WAIT: the `ready` word of the lane's ticket; L_SEND (group leader): the reserved position
    for (;;) {
      ++iter;
      if ((iter & 4095u) == 0u) { // watchdog: a frame never runs this long (decided by lane 0: the warp leaves together)
        const unsigned late = __shfl_sync(FULL, global_ns() - t_start > 4 * pp.timeout_ns ? 1u : 0u, 0);
        if (late) {
          if (lane == 0) atomicOr(&ctl->err, 1u);
          break;
        }
      }
      if (COUNT) pc0 = clock64();
      if (__any_sync(FULL, pend0 != nullptr)) publish(); // what the previous iteration stored: its payload has arrived by now
      // ---- SendRays, second half: the slots reserved in other ranks' inboxes one iteration ago (see below) have come back
      if (__any_sync(FULL, st == L_SEND)) {
        const unsigned base = __shfl_sync(FULL, landed, (team >> 8) & 31);
        if (st == L_SEND) {
          const int dest = team & 0xFF;
          const unsigned pos = base + ((unsigned)team >> 16);
          const RayQueue q = pp.peer_inbox[dest];
          if (pos < q.cap) {
            Ray r;
            reload(r);
            queue_store(q, pos, r, inst);
            pend_add(q.ready + pos, true);
          } else { // never silently: the frame fails with GVTB200_ERR_CAPACITY; the count stays balanced so that it still ends
            c_qovf++;
            atomicOr(&ctl->err, 2u);
            atomicAdd_system(&pp.mail.m[dest]->pc.consumed, 1u);
          }
          dirty = true;
          c_exp++, c_cons++;
          st = L_IDLE, team = -1;
        }
      }
      // ---- idle lanes draw inbox tickets ---------------------------------------------------------------------------
      const unsigned idle = __ballot_sync(FULL, st == L_IDLE);
      const unsigned tracing = __ballot_sync(FULL, st == L_ACTIVE || st == L_HELP);
      const bool starved = 32 - __popc(tracing) >= tune.exit_at; // the traversal loop would come straight back
      // (a ticket is drawn without looking whether its ray exists yet: a lane keeps its ticket until the ray lands, so at most
      // one ticket per lane is ever ahead of `tail`, arriving rays find a lane waiting, and a refill costs one L2 round trip)
      if (idle != 0u && (tracing == 0u || starved || __popc(idle) >= tune.refill_at)) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&ctl->head, (unsigned)__popc(idle));
        base = __shfl_sync(FULL, base, 0);
        if (st == L_IDLE) idx = base + (unsigned)__popc(idle & ltmask), st = L_WAIT;
      }
      // ---- lanes that hold a ticket: has the ray landed?  While other lanes of the warp trace, the question is asked now
      // and the answer is looked at after the traversal burst (an L2 round trip the whole warp would otherwise wait for)
      landed = 0u;
      if (st == L_WAIT) {
        if (idx >= pp.inbox.cap) st = L_DEAD; // beyond the inbox: no ray will ever land there (a producer that finds the inbox full flags the frame)
        else landed = ld_volatile_u32(pp.inbox.ready + idx);
      }
      auto start_landed = [&]() {
        if (st == L_WAIT && landed == seq) {
          const float4 c0 = ld_strong_f4(pp.inbox.c0 + idx), c1 = ld_strong_f4(pp.inbox.c1 + idx);
          const int4 c3 = ld_strong_i4(pp.inbox.c3 + idx);
          inst = ld_strong_i32(pp.inbox.inst + idx);
          type = c3.w;
          if (type == RAY_CAMERA_REF) { // a camera ray of another rank's film shard, sent as (index, routing distance): the
            Ray r;                      // record is rebuilt here, bit for bit what the sender had, and written into the slot
            camera_ray(fp.cam, (size_t)(unsigned)c3.x, r);
            r.o = xadd(fp.cam.eye, xmul(r.d, xmul(u2f((uint32_t)c3.y), 0.95f))); // TracerBase.h:394
            queue_store(pp.inbox, idx, r, inst);
            o = r.o, t_min = r.t_min, d = r.d, t_max = r.t_max, type = RAY_PRIMARY;
          } else {
            o = v3(c0.x, c0.y, c0.z), t_min = c0.w, d = v3(c1.x, c1.y, c1.z), t_max = c1.w;
          }
          any = (type == RAY_SHADOW), first = true, st = L_ACTIVE;
          begin_instance();
        }
      };
      if (tracing == 0u) start_landed(); // nothing to overlap with
      // ---- nothing to trace in this warp ----------------------------------------------------------------------------------
      if (__ballot_sync(FULL, st == L_ACTIVE || st == L_HELP || st == L_JOIN) == 0u) {
        flush();
        unsigned stop = 0;
        if (lane == 0) stop = ld_volatile_u32(&ctl->stop) == seq ? 1u : 0u;
        if (__shfl_sync(FULL, stop, 0)) break;
        __nanosleep(256);
        if ((iter & 255u) == 0u) { // a peer died, or a bug: never spin for ever
          const unsigned long long now = __shfl_sync(FULL, global_ns(), 0);
          if (t_idle0 == 0) t_idle0 = now;
          if (now - t_idle0 > 2 * pp.timeout_ns) {
            if (lane == 0) atomicOr(&ctl->err, 1u);
            break;
          }
        }
        if (COUNT) p_wait += clock64() - pc0;
        continue;
      }
      t_idle0 = 0;
      long long pc1 = 0;
      if (COUNT) pc1 = clock64(), p_poll += pc1 - pc0;
      // ---- traverse in lock step (comes back after tune.persist_trips iterations at the latest: tickets get polled and new
      // arrivals get lanes; whether the inbox holds rays is not known without asking, so it never leaves early for a refill)
      // ---- a nearly empty warp (the tail of the frame: a few long rays, each a chain of hundreds of dependent node visits)
      // shares its rays out: a free lane takes the bottom half of a traversing lane's stack -- the oldest, largest subtrees
      // -- and traverses it as a HELPER with a copy of the ray; hits are merged into the owner when a helper is done (smaller
      // t, then smaller primitive id: the result does not depend on who found what; any hit ends a shadow ray for the whole
      // team).  The chain of a long ray shrinks with the number of lanes working on it; nothing else changes: helpers never
      // touch the frame's counters, the owner finishes the ray when its team is done.
      {
        const unsigned tm0 = __ballot_sync(FULL, st == L_ACTIVE || st == L_HELP);
        if (tune.split_below > 0 && __popc(tm0) <= tune.split_below) {
          const unsigned donors = __ballot_sync(FULL, (st == L_ACTIVE || st == L_HELP) && tv.sp >= 2);
          const unsigned frees = __ballot_sync(FULL, st == L_IDLE || st == L_DEAD || (st == L_WAIT && landed != seq));
          const int n = min(__popc(donors), __popc(frees));
          if (n > 0) {
            const bool isD = ((donors >> lane) & 1u) && __popc(donors & ltmask) < n;
            const bool isF = ((frees >> lane) & 1u) && __popc(frees & ltmask) < n;
            // partner lane: the r-th donor pairs with the r-th free lane
            int partner = (int)lane;
            if (isD) partner = (int)__fns(frees, 0, __popc(donors & ltmask) + 1);
            if (isF) partner = (int)__fns(donors, 0, __popc(frees & ltmask) + 1);
            const int src = isF ? partner : (int)lane;
            // the ray, as the donor holds it
            const float ox = __shfl_sync(FULL, rl.o.x, src), oy = __shfl_sync(FULL, rl.o.y, src), oz = __shfl_sync(FULL, rl.o.z, src);
            const float dx = __shfl_sync(FULL, rl.d.x, src), dy = __shfl_sync(FULL, rl.d.y, src), dz = __shfl_sync(FULL, rl.d.z, src);
            const float rx = __shfl_sync(FULL, rl.rdx, src), ry = __shfl_sync(FULL, rl.rdy, src), rz = __shfl_sync(FULL, rl.rdz, src);
            const unsigned oc = __shfl_sync(FULL, rl.octinv, src);
            const unsigned long long np = __shfl_sync(FULL, (unsigned long long)nodes, src), tp = __shfl_sync(FULL, (unsigned long long)tris, src);
            const float tf = __shfl_sync(FULL, tfar, src), tmx = __shfl_sync(FULL, tv.tmax, src), ht = __shfl_sync(FULL, tv.hit.t, src);
            const int an = __shfl_sync(FULL, (int)any, src), dsp = __shfl_sync(FULL, tv.sp, src);
            const int own = __shfl_sync(FULL, st == L_HELP ? (team & 0xFF) : (int)lane, src);
            const int h = min(dsp / 2, TRACE_SM_STACK); // entries the helper takes: the bottom of the donor's stack (all in shared memory)
            if (isF) {
              const uint2 *dcol = s_stack + (threadIdx.x - lane + (unsigned)partner);
              for (int k = 0; k < h; ++k) stack.s[k * TRACE_THREADS] = dcol[k * TRACE_THREADS];
            }
            __syncwarp();
            if (isD) {
              const int hh = min(tv.sp / 2, TRACE_SM_STACK);
              for (int k = hh; k < tv.sp; ++k) stack.set(k - hh, stack.get(k));
              tv.sp -= hh;
            }
            if (isF) {
              rl.o = v3(ox, oy, oz), rl.d = v3(dx, dy, dz), rl.rdx = rx, rl.rdy = ry, rl.rdz = rz, rl.octinv = oc;
              nodes = (const float4 *)np, tris = (const float4 *)tp, tfar = tf, any = an != 0;
              tv.sp = h, tv.tmax = tmx;
              tv.ng = make_uint2(0u, 0u), tv.tg = make_uint2(GVT_TRI_TAG, 0u);
              tv.hit.prim = -1, tv.hit.slot = -1, tv.hit.t = ht, tv.hit.u = 0.f, tv.hit.v = 0.f;
              team = own | (st << 8); // what the lane was doing comes back when it is done helping
              st = L_HELP;
            }
            __syncwarp();
          }
        }
      }
      const unsigned resting = __ballot_sync(FULL, st != L_ACTIVE && st != L_HELP);
      const bool done = warp_traverse<COUNT, true>(st == L_ACTIVE || st == L_HELP, resting, false, tune.persist_trips, tune, nodes, tris, rl, tfar, any, tv, stack, tc, s_lut);
      long long pc2 = 0;
      if (COUNT) pc2 = clock64(), p_trav += pc2 - pc1;
      // ---- helpers that are done hand their hit to the ray's owner; owners wait for their team --------------------------------
      bool fin = st == L_ACTIVE && done;
      if (__ballot_sync(FULL, st == L_HELP || st == L_JOIN) != 0u) {
        unsigned hm = __ballot_sync(FULL, st == L_HELP && done);
        while (hm) {
          const int f = __ffs((int)hm) - 1;
          hm &= hm - 1u;
          const int own = __shfl_sync(FULL, team & 0xFF, f);
          const int hp = __shfl_sync(FULL, tv.hit.prim, f), hs = __shfl_sync(FULL, tv.hit.slot, f);
          const float ht = __shfl_sync(FULL, tv.hit.t, f), hu = __shfl_sync(FULL, tv.hit.u, f), hv = __shfl_sync(FULL, tv.hit.v, f);
          if ((int)lane == own && hp >= 0 && (ht < tv.hit.t || (ht == tv.hit.t && (tv.hit.prim < 0 || hp < tv.hit.prim)))) {
            tv.hit.prim = hp, tv.hit.slot = hs, tv.hit.t = ht, tv.hit.u = hu, tv.hit.v = hv;
            tv.tmax = ht;
          }
          if ((int)lane == f) st = team >> 8, team = -1;
        }
        // a shadow ray with a hit is over for its whole team
        const unsigned over = __ballot_sync(FULL, (st == L_ACTIVE || st == L_JOIN) && any && tv.hit.prim >= 0);
        if (st == L_HELP && ((over >> (team & 0xFF)) & 1u)) st = team >> 8, team = -1;
        if (st == L_ACTIVE && ((over >> lane) & 1u)) fin = true;
        // owners whose own stack is empty wait for their helpers
        unsigned helpers = __ballot_sync(FULL, st == L_HELP);
        bool has_helpers = false;
        while (helpers) {
          const int f = __ffs((int)helpers) - 1;
          helpers &= helpers - 1u;
          if (__shfl_sync(FULL, team & 0xFF, f) == (int)lane) has_helpers = true;
        }
        if (fin && has_helpers) st = L_JOIN, fin = false;
        else if (st == L_JOIN && !has_helpers) st = L_ACTIVE, fin = true;
      }
      // ---- rays that finished their instance -------------------------------------------------------------------------
      if (fin) {
        if (COUNT && pp.inst_load) atomicAdd(pp.inst_load + inst, (unsigned long long)(tc.nodes - nodes0));
        if (tv.hit.prim >= 0) {
          if (type != RAY_SHADOW) { // closest hit -> the shader SMs; an occluded shadow ray is dropped (:487-489)
            Ray r;
            reload(r);
            r.t = tv.hit.t;
            const unsigned pos = queue_reserve(&ctl->hit_tail);
            if (pos < pp.hitq.cap) {
              queue_store(pp.hitq, pos, r, inst);
              pp.hit[pos] = make_int4(tv.hit.prim, tv.hit.slot, (int)f2u(tv.hit.u), (int)f2u(tv.hit.v));
              pend_add(pp.hitq.ready + pos, false);
            } else {
              c_qovf++;
              atomicOr(&ctl->err, 2u);
              atomicAdd(&ctl->consumed, 1u);
            }
          }
          st = L_IDLE, c_cons++;
        } else {
          if (any && first) c_shok++;
          first = false;
          float t;
          const int next = tlas_route(sd.tlas, o, d, t_max, inst, false, t);
          if (next < 0) {
            if (type == RAY_SHADOW) { // TracerBase.h:396-399, straight into the root's image
              Ray r;
              reload(r);
              if (xlength(r.color) > 0.f) {
                const V3 c = xmul(r.color, r.w);
                if (me == pp.root) {
                  fb_add(pp.root_fb, r.id, c);
                } else {
                  float *px = pp.root_fb + 4 * (size_t)r.id;
                  atomicAdd_system(px, c.x), atomicAdd_system(px + 1, c.y), atomicAdd_system(px + 2, c.z), atomicAdd_system(px + 3, 1.0f);
                  dirty = true;
                }
                c_fb++;
              }
            }
            st = L_IDLE, c_cons++;
          } else {
            o = xadd(o, xmul(d, xmul(t, 0.95f))); // TracerBase.h:394
            if (to_light && type == RAY_SHADOW) t_max = xsub(t_max, xmul(t, 0.95f));
            const int dest = multi ? pick_server(pp.servers, next, sd.inst[next].owner, d) : me;
            if (dest == me) {
              inst = next;
              begin_instance();
            } else {
              // SendRays, first half: the slot in the destination's inbox is reserved now -- one system-scope atomic per
              // destination and warp, an NVLink round trip -- and used at the top of the next iteration: the answer
              // travels while the warp traverses instead of stalling it here
              cg::coalesced_group grp = cg::labeled_partition(cg::coalesced_threads(), dest);
              const unsigned leader = grp.shfl(lane, 0);
              if (grp.thread_rank() == 0) landed = atomicAdd_system(&pp.mail.m[dest]->pc.tail, grp.size());
              team = dest | (int)(leader << 8) | (int)(grp.thread_rank() << 16);
              inst = next;
              st = L_SEND;
            }
          }
        }
      }
      start_landed();
      if ((iter & 15u) == 0u) flush();
      if (COUNT) p_post += clock64() - pc2;
    }
    publish();
    flush();
    if (COUNT && lane == 0) {
      atomicAdd(&ctl->prof[1], (unsigned long long)p_wait), atomicAdd(&ctl->prof[2], (unsigned long long)p_trav);
      atomicAdd(&ctl->prof[3], (unsigned long long)p_post), atomicAdd(&ctl->prof[4], (unsigned long long)p_poll), atomicAdd(&ctl->prof[9], 1ull);
    }
    if (stack.overflow) c_ovf = 1;
    if (COUNT) {
      count_add(&sd.ctr->nodes_visited, tc.nodes);
      count_add(&sd.ctr->tris_tested, tc.tris);
    }
  }
  count_add(&sd.ctr->instance_visits, c_visits);
  count_add(&sd.ctr->shadow_unoccluded, c_shok);
  count_add(&sd.ctr->fb_adds, c_fb);
  count_add(&sd.ctr->rays_exported, c_exp);
  count_add(&sd.ctr->stack_overflows, c_ovf);
  count_add(&sd.ctr->queue_overflows, c_qovf);
  count_add(&sd.ctr->routed_initial, c_routed);
  count_add(&sd.ctr->primary_hits, c_hits);
  count_add(&sd.ctr->shadow_generated, c_shgen);
  count_add(&sd.ctr->secondary_generated, c_sec);
  if (blockIdx.x == 0 && threadIdx.x == 32) atomicAdd(&sd.ctr->camera_rays, (unsigned long long)pp.cam_count);
  // ---- the image is complete in the root's memory: every contribution was ordered before the `consumed` count the
  // root's control warp saw.  IceTComposite::localAdd's running clamp, applied once (min(1, sum), non-negative terms).
  if (pp.clamp && me == pp.root) {
    float4 *fb4 = reinterpret_cast<float4 *>(pp.root_fb);
    const size_t gtid = blockIdx.x * (size_t)TRACE_THREADS + threadIdx.x;
    for (size_t i = gtid; i < pp.fb_pix; i += (size_t)gridDim.x * TRACE_THREADS) {
      float4 v = __ldcg(fb4 + i);
      if (v.w != 0.f) {
        v.x = fminf(v.x, 1.f), v.y = fminf(v.y, 1.f), v.z = fminf(v.z, 1.f);
        fb4[i] = v;
      }
    }
  }
}

} // namespace gvt
