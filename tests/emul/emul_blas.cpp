// emul_blas.cpp -- TEST INFRASTRUCTURE.  Compiles the product's per-thread device bodies (gravit_b200/csrc/*.h)
// with g++ (GVT_HD = inline, -ffp-contract=off) and drives them sequentially, so that the BLAS build LOGIC
// (Morton keys, Karras tree, refit, wide collapse) and the compressed-wide-BVH traversal can be checked against
// the oracle on a box without a GPU.  Never linked into libgvtb200.so; the product has no CPU path.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../gravit_b200/csrc/bvh_build.h"
#include "../../gravit_b200/csrc/gvt_device.h"
#include "../../gravit_b200/csrc/shade.h"
#include "../../gravit_b200/csrc/traverse.h"

using namespace gvt;

struct EmulMesh {
  std::vector<float> verts;
  std::vector<uint32_t> faces;
  std::vector<uint32_t> keys_unsorted, keys, vals, morton_order;
  std::vector<float4> bn_lo, bn_hi;
  std::vector<int2> bn_child, bn_range;
  std::vector<int> bn_parent;
  std::vector<float4> wnodes, wtris;
  unsigned n_wnodes = 0, n_wtris = 0, depth = 0, bdepth = 0;
  Blas blas;
};


// The product's insertion-based optimisation (bvh_build.h reinsert_*), phases in the order the device runs them; every
// phase is a loop over nodes here and a kernel there.  The lock winners do not depend on the order of the bids, so the
// result is the device's.
static void emul_refit(BuildCtx &c) {
  const int n = (int)c.n;
  std::vector<int> stack, order;
  stack.push_back(0);
  while (!stack.empty()) {
    int b = stack.back();
    stack.pop_back();
    order.push_back(b);
    int2 ch = c.bn_child[b];
    if (ch.x < n - 1) stack.push_back(ch.x);
    if (ch.y < n - 1) stack.push_back(ch.y);
  }
  for (size_t i = order.size(); i-- > 0;) {
    int b = order[i];
    int2 ch = c.bn_child[b];
    float4 l0 = c.bn_lo[ch.x], h0 = c.bn_hi[ch.x], l1 = c.bn_lo[ch.y], h1 = c.bn_hi[ch.y];
    c.bn_lo[b] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
    c.bn_hi[b] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
    if (c.ri_cnt) relayout_count_node(c, b);
  }
}
static void reinsert_optimise(BuildCtx &c, std::vector<uint32_t> &vals, int iterations) {
  const int n = (int)c.n, total = 2 * n - 1;
  if (n < 4 || iterations <= 0) return;
  std::vector<int4> pick(total);
  std::vector<unsigned long long> lock(total);
  std::vector<unsigned> cnt(total, 0), go(total, 0);
  c.ri_pick = pick.data(), c.ri_lock = lock.data(), c.ri_go = go.data(), c.ri_cnt = nullptr;
  for (int it = 0; it < iterations; ++it) {
    std::fill(lock.begin(), lock.end(), 0ull);
    for (int x = 1; x < total; ++x) reinsert_search(c, x);
    pick[0] = make_int4(-1, -1, 0, 0);
    for (int x = 1; x < total; ++x) reinsert_lock(c, x);
    for (int x = 1; x < total; ++x) reinsert_check(c, x);
    go[0] = 0;
    for (int x = 1; x < total; ++x) reinsert_confirm(c, x);
    size_t moved = 0;
    for (int x = 1; x < total; ++x) moved += reinsert_apply(c, x) ? 1 : 0;
    emul_refit(c);
    if (getenv("GVT_EMUL_VERBOSE")) {
      double sah = 0;
      for (int b = 0; b < n - 1; ++b) sah += box_area4(c.bn_lo[b], c.bn_hi[b]);
      fprintf(stderr, "reinsert iteration %d: %zu moves, inner-node SAH %.3f\n", it, moved, sah / box_area4(c.bn_lo[0], c.bn_hi[0]));
    }
    if (!moved) break;
  }
  // relayout: leaves renumbered in depth-first order
  c.ri_cnt = cnt.data();
  emul_refit(c);
  std::vector<uint32_t> nvals(n);
  std::vector<unsigned> npos(n);
  std::vector<int> lpar(n), lslot(n);
  for (int k = 0; k < n; ++k) {
    const int id = n - 1 + k;
    npos[k] = relayout_first(c, id), lpar[k] = c.bn_parent[id], lslot[k] = c.bn_child[lpar[k]].y == id ? 1 : 0;
  }
  for (int b = 0; b < n - 1; ++b) {
    const unsigned f = relayout_first(c, b);
    c.bn_range[b] = make_int2((int)f, (int)(f + cnt[b] - 1));
  }
  for (int k = 0; k < n; ++k) {
    nvals[npos[k]] = vals[k];
    const int nid = n - 1 + (int)npos[k];
    c.bn_parent[nid] = lpar[k];
    if (lslot[k]) c.bn_child[lpar[k]].y = nid; else c.bn_child[lpar[k]].x = nid;
  }
  vals = nvals;
  c.vals = vals.data();
  for (int k = 0; k < n; ++k) leaf_setup(c, (unsigned)k);
  c.ri_pick = nullptr, c.ri_lock = nullptr, c.ri_go = nullptr, c.ri_cnt = nullptr;
}

#define EM_API extern "C" __attribute__((visibility("default")))

EM_API void *emul_build(const float *verts, unsigned nv, const uint32_t *faces, unsigned nf) {
  EmulMesh *m = new EmulMesh();
  m->verts.assign(verts, verts + 3 * (size_t)nv);
  m->faces.assign(faces, faces + 3 * (size_t)nf);
  BuildCtx c;
  memset(&c, 0, sizeof(c));
  c.verts = m->verts.data();
  c.faces = m->faces.data();
  c.n = nf;
  V3 lo = v3(FLT_MAX, FLT_MAX, FLT_MAX), hi = v3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  for (unsigned i = 0; i < nv; ++i) {
    lo.x = fastmin(lo.x, verts[3 * i]), lo.y = fastmin(lo.y, verts[3 * i + 1]), lo.z = fastmin(lo.z, verts[3 * i + 2]);
    hi.x = fastmax(hi.x, verts[3 * i]), hi.y = fastmax(hi.y, verts[3 * i + 1]), hi.z = fastmax(hi.z, verts[3 * i + 2]);
  }
  c.mlo = lo;
  c.mext = xsub(hi, lo);
  c.pad = blas_leaf_pad(c.mext);
  m->keys.resize(nf), m->vals.resize(nf);
  c.keys = m->keys.data(), c.vals = m->vals.data();
  for (unsigned f = 0; f < nf; ++f) tri_setup(c, f);
  m->keys_unsorted = m->keys;
  { // stable sort by key (what the LSD radix sort does)
    std::vector<uint32_t> order(nf);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return m->keys_unsorted[a] < m->keys_unsorted[b]; });
    for (unsigned i = 0; i < nf; ++i) m->keys[i] = m->keys_unsorted[order[i]], m->vals[i] = order[i];
  }
  size_t nb = 2 * (size_t)nf;
  m->bn_lo.resize(nb), m->bn_hi.resize(nb), m->bn_child.resize(nb), m->bn_range.resize(nb), m->bn_parent.assign(nb, -1);
  c.bn_lo = m->bn_lo.data(), c.bn_hi = m->bn_hi.data(), c.bn_child = m->bn_child.data(), c.bn_range = m->bn_range.data();
  c.bn_parent = m->bn_parent.data();
  m->morton_order = m->vals; // the Morton-sorted face order (parity hook); the SAH build re-orders vals
  const char *bmode = getenv("GVTB200_BUILD");
  const bool use_sah = !(bmode && !strcmp(bmode, "lbvh")) && nf > 1;
  if (use_sah) {
    // top-down binned SAH with the product's per-node bodies (bvh_build.h); the device runs the same bodies level by level
    std::vector<float4> tb_lo(nf), tb_hi(nf);
    for (unsigned f = 0; f < nf; ++f) tri_box_padded(c, f, tb_lo[f], tb_hi[f]);
    std::vector<uint32_t> idx(m->vals);
    std::vector<SahBin> bins(3 * GVT_SAH_BINS);
    int next_internal = 1;
    std::vector<SahNode> work;
    SahNode root;
    memset(&root, 0, sizeof(root));
    root.lo = 0, root.hi = (int)nf, root.id = 0;
    c.bn_range[0] = make_int2(0, (int)nf - 1);
    c.bn_parent[0] = -1;
    work.push_back(root);
    while (!work.empty()) {
      SahNode nd = work.back();
      work.pop_back();
      const int cnt = nd.hi - nd.lo;
      for (int q = 0; q < 3; ++q) nd.cb[q] = 0x7FFFFFFF, nd.cb[3 + q] = (int)0x80000000;
      for (int p = nd.lo; p < nd.hi; ++p) {
        float4 l = tb_lo[idx[p]], h = tb_hi[idx[p]];
        float cc[3] = { 0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z) };
        for (int q = 0; q < 3; ++q) nd.cb[q] = std::min(nd.cb[q], fenc(cc[q])), nd.cb[3 + q] = std::max(nd.cb[3 + q], fenc(cc[q]));
      }
      float lo[3], hi[3];
      if (cnt > GVT_SAH_SMALL) {
        for (auto &b : bins) {
          for (int q = 0; q < 3; ++q) b.lo[q] = 0x7FFFFFFF, b.hi[q] = (int)0x80000000;
          b.cnt = 0, b.pad = 0;
        }
        for (int p = nd.lo; p < nd.hi; ++p) {
          float4 l = tb_lo[idx[p]], h = tb_hi[idx[p]];
          for (int ax = 0; ax < 3; ++ax) {
            float cc = 0.5f * ((&l.x)[ax] + (&h.x)[ax]);
            SahBin &B = bins[ax * GVT_SAH_BINS + sah_bin(cc, fdec(nd.cb[ax]), fdec(nd.cb[3 + ax]))];
            B.lo[0] = std::min(B.lo[0], fenc(l.x)), B.lo[1] = std::min(B.lo[1], fenc(l.y)), B.lo[2] = std::min(B.lo[2], fenc(l.z));
            B.hi[0] = std::max(B.hi[0], fenc(h.x)), B.hi[1] = std::max(B.hi[1], fenc(h.y)), B.hi[2] = std::max(B.hi[2], fenc(h.z));
            B.cnt++;
          }
        }
        sah_eval_bins(bins.data(), nd, lo, hi);
      } else {
        sah_eval_small(tb_lo.data(), tb_hi.data(), idx.data(), nd, lo, hi);
      }
      c.bn_lo[nd.id] = make_float4(lo[0], lo[1], lo[2], 0.f);
      c.bn_hi[nd.id] = make_float4(hi[0], hi[1], hi[2], 0.f);
      std::vector<uint32_t> L, R;
      for (int p = nd.lo; p < nd.hi; ++p) (sah_goes_left(nd, tb_lo.data(), tb_hi.data(), idx[p], p) ? L : R).push_back(idx[p]);
      if (L.empty() || R.empty()) { // same guard as k_sah_children
        L.assign(idx.begin() + nd.lo, idx.begin() + nd.lo + cnt / 2);
        R.assign(idx.begin() + nd.lo + cnt / 2, idx.begin() + nd.hi);
      }
      std::copy(L.begin(), L.end(), idx.begin() + nd.lo);
      std::copy(R.begin(), R.end(), idx.begin() + nd.lo + L.size());
      int child[2];
      for (int side = 0; side < 2; ++side) {
        int clo = side ? nd.lo + (int)L.size() : nd.lo, chi = side ? nd.hi : nd.lo + (int)L.size();
        if (chi - clo == 1) {
          child[side] = (int)nf - 1 + clo;
        } else {
          SahNode ch;
          memset(&ch, 0, sizeof(ch));
          ch.lo = clo, ch.hi = chi, ch.id = next_internal++;
          c.bn_range[ch.id] = make_int2(clo, chi - 1);
          child[side] = ch.id;
          work.push_back(ch);
        }
        c.bn_parent[child[side]] = nd.id;
      }
      c.bn_child[nd.id] = make_int2(child[0], child[1]);
    }
    m->vals = idx;
    c.vals = m->vals.data();
    for (unsigned k = 0; k < nf; ++k) leaf_setup(c, k);
    reinsert_optimise(c, m->vals, reinsert_iterations(nf));
  } else {
  for (unsigned k = 0; k < nf; ++k) leaf_setup(c, k);
  for (unsigned i = 0; i + 1 < nf; ++i) karras_node(c, (int)i);

  // refit: post-order
  if (nf > 1) {
    std::vector<int> stack, order;
    stack.push_back(0);
    while (!stack.empty()) {
      int b = stack.back();
      stack.pop_back();
      order.push_back(b);
      int2 ch = c.bn_child[b];
      if (ch.x < (int)nf - 1) stack.push_back(ch.x);
      if (ch.y < (int)nf - 1) stack.push_back(ch.y);
    }
    for (size_t i = order.size(); i-- > 0;) {
      int b = order[i];
      int2 ch = c.bn_child[b];
      float4 l0 = c.bn_lo[ch.x], h0 = c.bn_hi[ch.x], l1 = c.bn_lo[ch.y], h1 = c.bn_hi[ch.y];
      c.bn_lo[b] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
      c.bn_hi[b] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
    }
  }
  }
  // optimal widening tables, bottom-up (children before parents)
  std::vector<float> dp_cost;
  std::vector<unsigned char> dp_pick;
  const char *wmode = getenv("GVTB200_WIDEN");
  if (nf > 1 && !(wmode && !strcmp(wmode, "greedy"))) {
    dp_cost.assign(7 * (size_t)nf, 0.f), dp_pick.assign(8 * (size_t)nf, 0);
    c.dp_cost = dp_cost.data(), c.dp_pick = dp_pick.data();
    std::vector<int> stack, order;
    stack.push_back(0);
    while (!stack.empty()) {
      int b = stack.back();
      stack.pop_back();
      order.push_back(b);
      int2 ch = c.bn_child[b];
      if (ch.x < (int)nf - 1) stack.push_back(ch.x);
      if (ch.y < (int)nf - 1) stack.push_back(ch.y);
    }
    for (size_t i = order.size(); i-- > 0;) collapse_dp_node(c, order[i]);
  }
  // collapse, level by level
  m->wnodes.assign(GVT_NODE_WORDS * (size_t)std::max(1u, nf), make_float4(0, 0, 0, 0));
  m->wtris.assign(3 * GVT_LEAF_MAX * (size_t)std::max(1u, nf), make_float4(0, 0, 0, 0));
  c.wnodes = m->wnodes.data(), c.wtris = m->wtris.data();
  unsigned node_count = 1, tri_count = 0, tri_slots = 0;
  if (nf > 0) {
    std::vector<CollapseTask> cur, next;
    CollapseTask root;
    root.bnode = (nf == 1) ? 0 /* leaf id n-1+0 == 0 */ : 0;
    root.wnode = 0;
    cur.push_back(root);
    while (!cur.empty()) {
      m->depth++;
      next.clear();
      for (auto &t : cur) {
        WideNode w;
        collapse_node(c, t, w);
        unsigned child_base = node_count, tri_base = tri_slots;
        node_count += w.n_internal, tri_slots += GVT_LEAF_MAX * w.n_leaves, tri_count += w.n_tris;
        size_t q = next.size();
        next.resize(q + w.n_internal);
        collapse_emit(c, t, w, child_base, tri_base, next.data() + q);
      }
      cur.swap(next);
    }
  }
  m->n_wnodes = node_count, m->n_wtris = tri_count;
  if (nf > 1) { // depth of the binary tree = levels of the level-synchronous device build
    std::vector<std::pair<int, int>> st2;
    st2.push_back({ 0, 1 });
    while (!st2.empty()) {
      auto [b, dep] = st2.back();
      st2.pop_back();
      m->bdepth = std::max(m->bdepth, (unsigned)dep);
      int2 ch = c.bn_child[b];
      if (ch.x < (int)nf - 1) st2.push_back({ ch.x, dep + 1 });
      if (ch.y < (int)nf - 1) st2.push_back({ ch.y, dep + 1 });
    }
  }
  memset(&m->blas, 0, sizeof(Blas));
  m->blas.nodes = m->wnodes.data();
  m->blas.tris = m->wtris.data();
  m->blas.n_tris = nf;
  m->blas.n_nodes = node_count;
  return m;
}
EM_API void emul_destroy(void *p) { delete (EmulMesh *)p; }
EM_API void emul_morton(void *p, uint32_t *codes, uint32_t *sorted_faces) {
  EmulMesh *m = (EmulMesh *)p;
  if (codes) memcpy(codes, m->keys_unsorted.data(), m->keys_unsorted.size() * 4);
  if (sorted_faces) memcpy(sorted_faces, m->morton_order.data(), m->morton_order.size() * 4);
}
EM_API void emul_stats(void *p, uint64_t *st) {
  EmulMesh *m = (EmulMesh *)p;
  st[0] = m->n_wnodes, st[1] = m->n_wtris, st[2] = m->depth, st[3] = m->bdepth;
}
struct Ray80 {
  float o[3], t_min, d[3], t_max, color[3], t;
  int id, depth;
  float w;
  int type;
  unsigned char pad[16];
};
struct HitRec {
  int prim;
  float t, u, v;
};
// closest hit (any = 0) or any hit (any = 1) of local-space rays; counts[0..1] += nodes visited / triangles tested
EM_API void emul_intersect(void *p, const void *rays, size_t n, int any, void *hits, uint64_t *counts) {
  EmulMesh *m = (EmulMesh *)p;
  const Ray80 *r = (const Ray80 *)rays;
  HitRec *h = (HitRec *)hits;
  TravCounters tc;
  tc.nodes = tc.tris = 0;
  uint64_t cn = 0, ct = 0;
  for (size_t i = 0; i < n; ++i) {
    Hit hit;
    V3 o = v3(r[i].o[0], r[i].o[1], r[i].o[2]), d = v3(r[i].d[0], r[i].d[1], r[i].d[2]);
    tc.nodes = tc.tris = 0;
    if (any)
      blas_traverse<true, true>(m->blas, o, d, GVT_RAY_EPSILON, FLT_MAX, hit, &tc);
    else
      blas_traverse<false, true>(m->blas, o, d, GVT_RAY_EPSILON, FLT_MAX, hit, &tc);
    cn += tc.nodes, ct += tc.tris;
    h[i].prim = hit.prim, h[i].t = hit.prim >= 0 ? hit.t : FLT_MAX, h[i].u = hit.u, h[i].v = hit.v;
  }
  if (counts) counts[0] += cn, counts[1] += ct;
}

// BVH-quality statistics (tools/bvh_eval.py): node visits split into bottom nodes (no internal child) and inner nodes,
// and a histogram of children per wide node.  out[0] inner visits, out[1] bottom visits, out[2] triangles tested,
// out[3..6] visits to the first 9 / 73 / 200 / 585 nodes (the top levels: nodes are numbered level by level)
EM_API void emul_visit_stats(void *p, const void *rays, size_t n, int any, uint64_t *out) {
  EmulMesh *m = (EmulMesh *)p;
  const Ray80 *r = (const Ray80 *)rays;
  for (size_t i = 0; i < n; ++i) {
    V3 o = v3(r[i].o[0], r[i].o[1], r[i].o[2]), d = v3(r[i].d[0], r[i].d[1], r[i].d[2]);
    RayLocal rl;
    ray_local(rl, o, d);
    Trav tv;
    trav_begin(tv, m->blas.n_tris, FLT_MAX);
    LocalStack stack;
    TravCounters tc;
    tc.nodes = tc.tris = 0;
    bool more = trav_has_nodes(tv);
    while (more) {
      if (trav_has_nodes(tv)) {
        { // index of the node about to be visited (what node_step computes): visits to the first K nodes -> out[3..6]
          const uint32_t hits8 = tv.ng.y >> 24, imask = tv.ng.y & 0xFFu, sl = best_slot(hits8, rl.octinv);
          const uint32_t ni = tv.ng.x + (uint32_t)popc32(imask & ((1u << sl) - 1u));
          const uint32_t K[4] = { 9u, 73u, 200u, 585u };
          for (int q = 0; q < 4; ++q) out[3 + q] += ni < K[q] ? 1u : 0u;
        }
        const unsigned before = tc.nodes;
        node_step<true, false>(m->blas.nodes, rl, GVT_RAY_EPSILON, tv, stack, &tc);
        if (tc.nodes != before) { // (a group dropped without a fetch is not a visit)
          out[(tv.ng.y & 0xFFu) ? 0 : 1]++;
          if ((tv.ng.y >> 24) == 0u && (tv.tg.y & 0x00FFFFFFu) == 0u) out[7]++; // nothing inside was hit: a stale or empty visit
        }
      }
      bool done = false;
      while (trav_has_tris(tv) && !done) done = tri_step<true>(m->blas.tris, rl, GVT_RAY_EPSILON, FLT_MAX, any != 0, tv, &tc);
      if (done) break;
      more = trav_next(tv, stack);
    }
    out[2] += tc.tris;
  }
}
// hist[k] = wide nodes with k children (k = 0..8); hist[9 + k] = leaf slots with k triangles (k = 1..3); hist[13] = bottom nodes
EM_API void emul_node_hist(void *p, uint64_t *hist) {
  EmulMesh *m = (EmulMesh *)p;
  for (unsigned i = 0; i < m->n_wnodes; ++i) {
    const uint4 *nd = reinterpret_cast<const uint4 *>(m->wnodes.data()) + GVT_NODE_WORDS * (size_t)i;
    uint32_t imask = nd[0].w >> 24, tmask = nd[1].z;
    int k = popc32(imask) + popc32(tmask & 0xFFu);
    hist[k]++;
    if (!imask) hist[13]++;
    for (int s = 0; s < 8; ++s)
      if ((tmask >> s) & 1u) hist[9 + ((tmask >> (8 + s)) & 1u) + ((tmask >> (16 + s)) & 1u) + 1]++;
  }
}

// EXPLORATION (tools/bvh_eval.py): node visits of a closest-hit traversal that always opens the globally nearest pending
// child and drops entries behind the closest hit found so far -- the lower bound for this BVH that the octant-ordered,
// distance-free stack of the product (traverse.h) is compared with.  Child entry distances are decoded exactly like node_step.
#include <queue>
EM_API void emul_ideal_visits(void *p, const void *rays, size_t n, uint64_t *out) {
  EmulMesh *m = (EmulMesh *)p;
  const Ray80 *r = (const Ray80 *)rays;
  const float4 *nodes = m->blas.nodes, *tris = m->blas.tris;
  for (size_t i = 0; i < n; ++i) {
    V3 o = v3(r[i].o[0], r[i].o[1], r[i].o[2]), d = v3(r[i].d[0], r[i].d[1], r[i].d[2]);
    RayLocal rl;
    ray_local(rl, o, d);
    float tmax = FLT_MAX;
    const float tnear = GVT_RAY_EPSILON;
    typedef std::pair<float, uint32_t> QE; // (-entry distance, node index)
    std::priority_queue<QE> pq;
    if (m->blas.n_tris) pq.push(QE(0.f, 0u));
    while (!pq.empty()) {
      QE e = pq.top();
      pq.pop();
      if (-e.first > tmax) break;
      out[0]++;
      const float4 *np = nodes + GVT_NODE_WORDS * (size_t)e.second;
      const uint32_t sx = (rl.octinv >> 2) & 1u, sy = (rl.octinv >> 1) & 1u, sz = rl.octinv & 1u;
      const float4 h0 = np[0], h1 = np[1];
      const float4 nx = np[3 - sx], fx = np[2 + sx], ny = np[5 - sy], fy = np[4 + sy], nz = np[7 - sz], fz = np[6 + sz];
      const uint32_t ew = f2u(h0.w);
      const float ax = u2f((ew & 0xFFu) << 23) * rl.rdx, ay = u2f(((ew >> 8) & 0xFFu) << 23) * rl.rdy, az = u2f(((ew >> 16) & 0xFFu) << 23) * rl.rdz;
      const float bx = fmaf(h0.x - rl.o.x, rl.rdx, -ax), by = fmaf(h0.y - rl.o.y, rl.rdy, -ay), bz = fmaf(h0.z - rl.o.z, rl.rdz, -az);
      const uint32_t imask = ew >> 24, tmask = f2u(h1.z);
      float tn[8], tf[8];
      for (int j = 0; j < 4; ++j) {
        const float *cnx = &nx.x, *cfx = &fx.x, *cny = &ny.x, *cfy = &fy.x, *cnz = &nz.x, *cfz = &fz.x;
        auto hi16 = [](float w) { return u2f(f2u(w) & 0xFFFF0000u); };
        auto lo16 = [](float w) { return u2f(f2u(w) << 16); };
        tn[j] = fmaxf(fmaxf(fmaf(hi16(cnx[j]), ax, bx), fmaf(hi16(cny[j]), ay, by)), fmaxf(fmaf(hi16(cnz[j]), az, bz), tnear));
        tf[j] = fminf(fminf(fmaf(hi16(cfx[j]), ax, bx), fmaf(hi16(cfy[j]), ay, by)), fminf(fmaf(hi16(cfz[j]), az, bz), tmax));
        tn[j + 4] = fmaxf(fmaxf(fmaf(lo16(cnx[j]), ax, bx), fmaf(lo16(cny[j]), ay, by)), fmaxf(fmaf(lo16(cnz[j]), az, bz), tnear));
        tf[j + 4] = fminf(fminf(fmaf(lo16(cfx[j]), ax, bx), fmaf(lo16(cfy[j]), ay, by)), fminf(fmaf(lo16(cfz[j]), az, bz), tmax));
      }
      const uint32_t child_base = f2u(h1.x), tri_base = f2u(h1.y);
      int rank_i = 0, rank_l = 0;
      for (int sl = 0; sl < 8; ++sl) {
        const bool hit = tn[sl] <= tf[sl];
        if ((imask >> sl) & 1u) {
          if (hit) pq.push(QE(-tn[sl], child_base + (uint32_t)rank_i));
          rank_i++;
        } else if ((tmask >> sl) & 1u) {
          if (hit)
            for (int k = 0; k < 3; ++k)
              if ((tmask >> (8 * k + sl)) & 1u) {
                const float4 *tp = tris + 3 * (size_t)(tri_base + 3u * (uint32_t)rank_l + (uint32_t)k);
                out[1]++;
                float t, u, v;
                if (tri_test(rl.o, rl.d, v3(tp[0].x, tp[0].y, tp[0].z), v3(tp[1].x, tp[1].y, tp[1].z), v3(tp[2].x, tp[2].y, tp[2].z), tnear, FLT_MAX, t, u, v) && t < tmax) tmax = t;
              }
          rank_l++;
        }
      }
    }
  }
}

// camera rays through the product's camera_ray body
EM_API void emul_camera_rays(const float *eye, const float *u, const float *v, const float *w, const float *scal7, int W, int H,
                             int samples, int depth, void *out) {
  Camera c;
  c.eye = v3(eye[0], eye[1], eye[2]), c.u = v3(u[0], u[1], u[2]), c.v = v3(v[0], v[1], v[2]), c.w = v3(w[0], w[1], w[2]);
  c.horz = scal7[0], c.vert = scal7[1], c.wmult = scal7[2], c.hmult = scal7[3], c.half_sample = scal7[4], c.offset = scal7[5],
  c.contri = scal7[6];
  c.W = W, c.H = H, c.samples = samples, c.depth = depth;
  Ray80 *o = (Ray80 *)out;
  size_t n = (size_t)W * H * samples * samples;
  for (size_t i = 0; i < n; ++i) {
    Ray r;
    camera_ray(c, i, r);
    memset(&o[i], 0, sizeof(Ray80));
    o[i].o[0] = r.o.x, o[i].o[1] = r.o.y, o[i].o[2] = r.o.z, o[i].t_min = r.t_min;
    o[i].d[0] = r.d.x, o[i].d[1] = r.d.y, o[i].d[2] = r.d.z, o[i].t_max = r.t_max;
    o[i].t = r.t, o[i].id = r.id, o[i].depth = r.depth, o[i].w = r.w, o[i].type = r.type;
  }
}
