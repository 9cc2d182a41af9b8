// pt_build.cpp — see pt_build.h.  Plain host C++ (no CUDA), compiled into libptcore.so.
#include "pt_build.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <chrono>
#include <thread>

namespace pt {
namespace {

inline const float *tri_ptr(const MeshBuild &m, int64_t i) { return m.tris.data() + i * 12; }

// ------------------------------------------------------------------------------------------------------------
// Step 1: the reference's median-split build, restated only as far as its shape matters for visibility.
// bvh.rs:15-76: bounds over the vertices; leaf when n <= 4 or depth >= 25; axis = x if strictly the largest
// extent, else y if larger than z, else z; sort the index slice by centroid ((v0+v1+v2) * (1/3)) on that axis;
// split at n/2.  Rust's sort_unstable_by leaves the order of equal keys unspecified (it depends on the std
// version); this restatement uses a STABLE sort, the same choice the oracle documents (tolerance class T6).
struct RefCtx {
  const MeshBuild *m;
  const float *cen;  // n x 3 centroids
  uint8_t *dead;
  std::atomic<int64_t> nodes{0}, leaves{0};
  std::atomic<int32_t> max_depth{0};
  int tie_kind = 0;
  uint64_t tie_seed = 0;
};

// The reference sorts a node's triangles with `sort_unstable_by` on one centroid coordinate (bvh.rs:45-53); the
// restatement uses a STABLE order (tolerance class T6).  For big nodes that order is produced by an LSD radix
// sort on the float's order-preserving integer image (three 11-bit passes over (key, index) pairs; -0.0 keyed as +0.0
// because the comparator calls them equal): O(n) per node instead of O(n log n) with an indirect comparator, which was
// 1.75 s of the 3.1 s commit of the 2 M-triangle scene.  PTC_REF_STABLE_SORT=1 forces std::stable_sort everywhere
// (tests compare the two).
inline uint32_t float_sort_key(float f) {
  if (f == 0.0f) f = 0.0f;  // -0.0 -> +0.0
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// Tie order of equal centroid keys (PTC_REF_TIE, read per build): what Rust's sort_unstable_by does with them is
// unspecified and changes with the std version (pdqsort up to 1.80, ipnsort since), and it decides which triangles end up
// together under a flat node.  The switch exists to MEASURE that sensitivity (tools/tie_order_study.py):
//   stable (default)  equal keys keep their input order
//   reverse           equal keys in reversed input order
//   random:<seed>     equal keys ordered by a hash of (seed, triangle index)
// The oracle (oracle/oracle.cpp: bvh_build) has the same switch; GPU-vs-oracle parity is tested under each.
struct TieMode {
  int kind = 0;  // 0 stable, 1 reverse, 2 random
  uint64_t seed = 0;
};
TieMode tie_mode_from_env() {
  TieMode t;
  const char *e = getenv("PTC_REF_TIE");
  if (!e || !*e || !strcmp(e, "stable")) return t;
  if (!strcmp(e, "reverse")) t.kind = 1;
  else if (!strncmp(e, "random:", 7)) t.kind = 2, t.seed = strtoull(e + 7, nullptr, 10);
  else throw std::invalid_argument(std::string("PTC_REF_TIE: unknown tie order '") + e + "' (stable | reverse | random:<seed>)");
  return t;
}
inline uint64_t tie_hash(uint64_t seed, uint64_t i) {  // splitmix64
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + i + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
void sort_by_centroid(int64_t *idx, int64_t n, const float *cen, int axis, uint64_t *scratch, const TieMode &tie) {
  static const bool force_std = getenv("PTC_REF_STABLE_SORT") != nullptr;
  if (tie.kind != 0) {
    // secondary key that realises the tie order, then one stable sort on (key, secondary)
    std::vector<std::pair<uint64_t, int64_t>> sec((size_t)n);
    for (int64_t i = 0; i < n; i++) sec[(size_t)i] = {tie.kind == 1 ? (uint64_t)(n - 1 - i) : tie_hash(tie.seed, (uint64_t)idx[i]), idx[i]};
    std::stable_sort(sec.begin(), sec.end(), [cen, axis](const std::pair<uint64_t, int64_t> &a, const std::pair<uint64_t, int64_t> &b) {
      const float ka = cen[a.second * 3 + axis], kb = cen[b.second * 3 + axis];
      if (ka < kb) return true;
      if (kb < ka) return false;
      return a.first < b.first;
    });
    for (int64_t i = 0; i < n; i++) idx[i] = sec[(size_t)i].second;
    return;
  }
  if (n < 4096 || force_std || scratch == nullptr) {
    std::stable_sort(idx, idx + n, [cen, axis](int64_t a, int64_t b) { return cen[a * 3 + axis] < cen[b * 3 + axis]; });
    return;
  }
  uint64_t *a = scratch, *b = scratch + n;
  for (int64_t i = 0; i < n; i++) a[i] = ((uint64_t)float_sort_key(cen[idx[i] * 3 + axis]) << 32) | (uint64_t)(uint32_t)idx[i];
  for (int pass = 0; pass < 3; pass++) {
    const int shift = 32 + 11 * pass;
    const uint64_t mask = pass == 2 ? 0x3ffu : 0x7ffu;
    uint32_t count[2049];
    memset(count, 0, sizeof(count));
    for (int64_t i = 0; i < n; i++) count[((a[i] >> shift) & mask) + 1]++;
    for (int k = 0; k < 2048; k++) count[k + 1] += count[k];
    for (int64_t i = 0; i < n; i++) b[count[(a[i] >> shift) & mask]++] = a[i];
    std::swap(a, b);
  }
  for (int64_t i = 0; i < n; i++) idx[i] = (int64_t)(uint32_t)a[i];
}

// `scratch`: 2 * n words for this call (children use disjoint halves), or nullptr
void ref_build(RefCtx &c, int64_t *idx, int64_t n, int depth, bool dead_in, int par_levels, uint64_t *scratch) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = 0; i < n; i++) {
    const float *t = tri_ptr(*c.m, idx[i]);
    for (int v = 0; v < 3; v++)
      for (int a = 0; a < 3; a++) {
        lo[a] = fminf(lo[a], t[v * 3 + a]);
        hi[a] = fmaxf(hi[a], t[v * 3 + a]);
      }
  }
  c.nodes.fetch_add(1, std::memory_order_relaxed);
  int32_t md = c.max_depth.load(std::memory_order_relaxed);
  while (depth > md && !c.max_depth.compare_exchange_weak(md, depth)) {
  }
  // aabb.rs:27-45: on a zero-extent axis t0 == t1, so `t_max <= t_min` holds and the node is never entered
  const bool flat = (lo[0] == hi[0]) || (lo[1] == hi[1]) || (lo[2] == hi[2]);
  const bool dead = dead_in || flat;
  if (n <= 4 || depth >= 25) {
    c.leaves.fetch_add(1, std::memory_order_relaxed);
    for (int64_t i = 0; i < n; i++) c.dead[idx[i]] = dead ? 1 : 0;
    return;
  }
  const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  const int axis = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);
  TieMode tie;
  tie.kind = c.tie_kind, tie.seed = c.tie_seed;
  sort_by_centroid(idx, n, c.cen, axis, scratch, tie);
  const int64_t mid = n / 2;
  uint64_t *s_left = scratch, *s_right = scratch ? scratch + 2 * mid : nullptr;
  if (par_levels > 0 && n > (1 << 14)) {
    std::thread th([&]() { ref_build(c, idx, mid, depth + 1, dead, par_levels - 1, s_left); });
    ref_build(c, idx + mid, n - mid, depth + 1, dead, par_levels - 1, s_right);
    th.join();
  } else {
    ref_build(c, idx, mid, depth + 1, dead, 0, s_left);
    ref_build(c, idx + mid, n - mid, depth + 1, dead, 0, s_right);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Step 2: binned SAH binary BVH over the live triangles.
struct Box3 {
  float lo[3], hi[3];
  void reset() {
    for (int a = 0; a < 3; a++) {
      lo[a] = INFINITY;
      hi[a] = -INFINITY;
    }
  }
  void grow(const Box3 &b) {
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], b.lo[a]);
      hi[a] = std::max(hi[a], b.hi[a]);
    }
  }
  void grow(const float *p) {
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], p[a]);
      hi[a] = std::max(hi[a], p[a]);
    }
  }
  double area() const {
    const double x = (double)hi[0] - lo[0], y = (double)hi[1] - lo[1], z = (double)hi[2] - lo[2];
    if (x < 0) return 0.0;
    return 2.0 * (x * y + y * z + z * x);
  }
};

struct B2 {
  Box3 box;
  int32_t left = -1, right = -1;  // left < 0: leaf
  int32_t first = 0, count = 0;   // range in the prim array
};

struct SahCtx {
  std::vector<B2> nodes;
  std::atomic<int32_t> next{0};
  const Box3 *pbox;   // per live prim
  const float *pcen;  // per live prim x 3
  int32_t *prim;      // permutation of live prim ids
};

constexpr int kBins = 16;
constexpr int kLeafMax = 3;  // a leaf slot of a wide node addresses at most 3 triangles

// `pbox` / `pcen` / `prim` are the arrays this subtree works on: the mesh-wide ones near the root, private compact copies
// below kLocalizeCount triangles (see there); `offset` turns positions in `prim` into positions in the mesh-wide list.
constexpr int32_t kLocalizeCount = 1 << 17;
int32_t sah_build_view(SahCtx &c, const Box3 *pbox, const float *pcen, int32_t *prim, int32_t offset, bool localized, int32_t first,
                       int32_t count, int par_levels) {
  if (!localized && count <= kLocalizeCount && count > 4096) {
    // From here down the subtree fits a core's cache — if its data are contiguous.  Near the root every pass gathers
    // boxes and centroids through the permutation (one DRAM miss per triangle, pass and level: what the build spent its
    // time on); a private copy in current order makes every deeper pass a stream.
    std::vector<Box3> lbox((size_t)count);
    std::vector<float> lcen((size_t)count * 3);
    std::vector<int32_t> lprim((size_t)count), orig((size_t)count);
    for (int32_t i = 0; i < count; i++) {
      const int32_t p = prim[first + i];
      orig[(size_t)i] = p;
      lbox[(size_t)i] = pbox[p];
      for (int a = 0; a < 3; a++) lcen[(size_t)i * 3 + a] = pcen[3 * (size_t)p + a];
      lprim[(size_t)i] = i;
    }
    const int32_t r = sah_build_view(c, lbox.data(), lcen.data(), lprim.data(), offset + first, true, 0, count, par_levels);
    for (int32_t i = 0; i < count; i++) prim[first + i] = orig[(size_t)lprim[(size_t)i]];
    return r;
  }
  const int32_t me = c.next.fetch_add(1);
  B2 node;
  node.first = first + offset;
  node.count = count;
  node.box.reset();
  Box3 cb;
  cb.reset();
  // Both passes over the node's triangles (bounds, then the bins of all three axes at once) are split over the threads
  // this subtree may still use while the node is large: the top three levels of a 2 M-triangle build used to run them on
  // one thread each, three gather passes per node.
  const int workers = (par_levels > 0 && count > (1 << 16)) ? std::min(1 << par_levels, 16) : 1;
  auto for_chunks = [&](auto &&body) {  // body(worker, begin, end)
    if (workers == 1) {
      body(0, first, first + count);
      return;
    }
    std::vector<std::thread> th;
    const int32_t per = (count + workers - 1) / workers;
    for (int w = 1; w < workers; w++) {
      const int32_t lo = first + w * per, hi = std::min(first + count, lo + per);
      if (lo < hi) th.emplace_back([&body, w, lo, hi]() { body(w, lo, hi); });
    }
    body(0, first, std::min(first + count, first + per));
    for (auto &t : th) t.join();
  };
  {
    Box3 nb1[2], *nb = nb1, *cbs = nb1 + 1;  // the common case (one worker) stays off the heap
    std::vector<Box3> nbv;
    if (workers > 1) nbv.resize(2 * (size_t)workers), nb = nbv.data(), cbs = nbv.data() + workers;
    for (int w = 0; w < workers; w++) nb[(size_t)w].reset(), cbs[(size_t)w].reset();
    for_chunks([&](int w, int32_t lo, int32_t hi) {
      Box3 x = nb[(size_t)w], y = cbs[(size_t)w];
      for (int32_t i = lo; i < hi; i++) {
        x.grow(pbox[prim[i]]);
        y.grow(pcen + 3 * (size_t)prim[i]);
      }
      nb[(size_t)w] = x, cbs[(size_t)w] = y;
    });
    for (int w = 0; w < workers; w++) {
      if (nb[(size_t)w].lo[0] <= nb[(size_t)w].hi[0]) node.box.grow(nb[(size_t)w]);
      if (cbs[(size_t)w].lo[0] <= cbs[(size_t)w].hi[0]) cb.grow(cbs[(size_t)w]);
    }
  }
  if (count <= kLeafMax) {
    c.nodes[me] = node;
    return me;
  }
  // A bin is initialised by the first triangle that falls into it (`used` = which bins hold anything): most of the
  // tree's ~N nodes are small, and resetting and sweeping 3 x 16 mostly empty bins for each of them was a third of the
  // build.  The sweep below visits the non-empty bins only; a split position behind an empty bin separates the same two
  // sets as the position before it, at the same cost, and was never chosen (strict <) — so the tree is the same, bit for bit.
  struct Bins {
    Box3 bb[3][kBins];
    int32_t bn[3][kBins];
    uint32_t used[3];
  };
  float kk[3];
  bool axis_ok[3];
  for (int a = 0; a < 3; a++) {
    const float cext = cb.hi[a] - cb.lo[a];
    axis_ok[a] = cext > 0.0f;
    kk[a] = axis_ok[a] ? (float)kBins / cext : 0.0f;
  }
  Bins wb1, *wb = &wb1;
  std::vector<Bins> wbv;
  if (workers > 1) wbv.resize((size_t)workers), wb = wbv.data();
  for (int w = 0; w < workers; w++) wb[(size_t)w].used[0] = wb[(size_t)w].used[1] = wb[(size_t)w].used[2] = 0u;
  for_chunks([&](int w, int32_t lo, int32_t hi) {
    Bins &B = wb[(size_t)w];
    uint32_t used[3] = {0u, 0u, 0u};
    for (int32_t i = lo; i < hi; i++) {
      const int32_t p = prim[i];
      const Box3 &pb = pbox[p];
      const float *pc = pcen + 3 * (size_t)p;
      for (int a = 0; a < 3; a++) {
        if (!axis_ok[a]) continue;
        int bi = (int)((pc[a] - cb.lo[a]) * kk[a]);
        bi = std::min(std::max(bi, 0), kBins - 1);
        if (used[a] >> bi & 1u) {
          B.bb[a][bi].grow(pb);
          B.bn[a][bi]++;
        } else {
          used[a] |= 1u << bi;
          B.bb[a][bi] = pb;
          B.bn[a][bi] = 1;
        }
      }
    }
    for (int a = 0; a < 3; a++) B.used[a] = used[a];
  });
  for (int w = 1; w < workers; w++)
    for (int a = 0; a < 3; a++)
      for (int bi = 0; bi < kBins; bi++) {
        if (!(wb[(size_t)w].used[a] >> bi & 1u)) continue;
        if (wb[0].used[a] >> bi & 1u) {
          wb[0].bb[a][bi].grow(wb[(size_t)w].bb[a][bi]);
          wb[0].bn[a][bi] += wb[(size_t)w].bn[a][bi];
        } else {
          wb[0].used[a] |= 1u << bi;
          wb[0].bb[a][bi] = wb[(size_t)w].bb[a][bi];
          wb[0].bn[a][bi] = wb[(size_t)w].bn[a][bi];
        }
      }
  int best_axis = -1, best_split = -1;
  double best_cost = DBL_MAX;
  for (int a = 0; a < 3; a++) {
    if (!axis_ok[a]) continue;
    const Box3 *bb = wb[0].bb[a];
    const int32_t *bn = wb[0].bn[a];
    int idx[kBins], m = 0;  // the non-empty bins, ascending
    for (uint32_t u = wb[0].used[a]; u != 0u; u &= u - 1u) idx[m++] = __builtin_ctz(u);
    if (m < 2) continue;  // everything in one bin: no position separates anything
    double right_area[kBins];  // [j]: bins idx[j] and above
    int32_t right_n[kBins];
    Box3 acc;
    acc.reset();
    int32_t cnt = 0;
    for (int j = m - 1; j > 0; j--) {
      acc.grow(bb[idx[j]]);
      cnt += bn[idx[j]];
      right_area[j] = acc.area();
      right_n[j] = cnt;
    }
    acc.reset();
    cnt = 0;
    for (int j = 0; j < m - 1; j++) {
      acc.grow(bb[idx[j]]);
      cnt += bn[idx[j]];
      const double cost = acc.area() * cnt + right_area[j + 1] * right_n[j + 1];
      if (cost < best_cost) {
        best_cost = cost;
        best_axis = a;
        best_split = idx[j] + 1;  // bins [0, best_split) go left
      }
    }
  }
  int32_t mid;
  if (best_axis < 0) {
    mid = first + count / 2;  // all centroids coincide
  } else {
    const float cext = cb.hi[best_axis] - cb.lo[best_axis];
    const float k = (float)kBins / cext;
    const float lo = cb.lo[best_axis];
    const int a = best_axis, split = best_split;
    int32_t *b = prim + first, *e = prim + first + count;
    int32_t *m = std::partition(b, e, [&](int32_t p) {
      int bin = (int)((pcen[3 * (size_t)p + a] - lo) * k);
      bin = std::min(std::max(bin, 0), kBins - 1);
      return bin < split;
    });
    mid = first + (int32_t)(m - b);
    if (mid == first || mid == first + count) mid = first + count / 2;
  }
  int32_t l, r;
  if (par_levels > 0 && count > (1 << 14)) {
    std::thread th([&]() { l = sah_build_view(c, pbox, pcen, prim, offset, localized, first, mid - first, par_levels - 1); });
    r = sah_build_view(c, pbox, pcen, prim, offset, localized, mid, first + count - mid, par_levels - 1);
    th.join();
  } else {
    l = sah_build_view(c, pbox, pcen, prim, offset, localized, first, mid - first, 0);
    r = sah_build_view(c, pbox, pcen, prim, offset, localized, mid, first + count - mid, 0);
  }
  node.left = l;
  node.right = r;
  c.nodes[me] = node;
  return me;
}

int32_t sah_build(SahCtx &c, int32_t first, int32_t count, int par_levels) {
  return sah_build_view(c, c.pbox, c.pcen, c.prim, 0, false, first, count, par_levels);
}

// ------------------------------------------------------------------------------------------------------------
// Step 3: 8-wide collapse + quantisation.
inline float u2f_host(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint32_t f2u_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

struct QFrame {
  float origin[3];
  int biased_exp[3];
  double scale[3];
};

// Frame = node box padded on every side, so that a child plane that coincides with the node's own plane still
// quantises with slack; the traversal's fused slab arithmetic then cannot cull a box whose face carries the hit.
QFrame make_frame(const Box3 &box) {
  QFrame f;
  for (int a = 0; a < 3; a++) {
    const float lo = box.lo[a], hi = box.hi[a];
    const float ext = hi - lo;
    const float maxabs = std::max(std::fabs(lo), std::fabs(hi));
    const float pad = ext * (1.0f / 512.0f) + maxabs * 0x1p-20f + 1e-30f;
    float flo = lo - pad, fhi = hi + pad;
    if (!(flo < lo)) flo = std::nextafter(lo, -INFINITY);
    if (!(fhi > hi)) fhi = std::nextafter(hi, INFINITY);
    const double need = ((double)fhi - (double)flo) / 255.0;
    int ex;
    std::frexp(need, &ex);  // need = m * 2^ex, m in [0.5, 1)  =>  2^ex >= need
    int biased = ex + 127;
    if (biased < 1) biased = 1;
    if (biased > 254 - 15) throw std::runtime_error("mesh extent too large for the quantised BVH frame");  // trav_node adds 15
    f.origin[a] = flo;
    f.biased_exp[a] = biased;
    f.scale[a] = std::ldexp(1.0, biased - 127);
  }
  return f;
}

// Which binary subtrees become the (up to eight) children of a wide node.  The greedy rule (open the child with the
// largest area until there are eight) leaves the bottom of the tree badly filled — 4.3 children per node and 1.8
// triangles per leaf on the 2 M-triangle mesh — and the traversal pays the full ~330 instructions of a node step
// whatever the fill.  The plan below is the dynamic programme of Ylitie, Karras & Laine 2017 (section 4.2) over the
// binary SAH tree:  C(n, i) = cheapest way to turn the subtree of n into at most i children of some wide node,
//     C(n, 1) = min( area(n) * tris(n) * c_t          if tris(n) <= 3   (the whole subtree as ONE leaf),
//                    area(n) * c_n + D(n, 8) )                          (an inner wide node with up to 8 children)
//     C(n, i) = min( D(n, i), C(n, i - 1) ),   D(n, j) = min_k C(left, k) + C(right, j - k).
struct CollapsePlan {
  std::vector<float> cost;     // [n * 8 + i - 1], i = 1..7
  std::vector<uint8_t> split;  // [n * 8 + i - 1]: triangles/children given to the LEFT subtree for D(n, i), 0 = "same as i - 1";
                               // entry i = 8 (index 7): the split of the node's own eight slots
  std::vector<uint8_t> leaf;   // C(n, 1) is realised as a leaf
  bool greedy = false;
};
CollapsePlan plan_collapse(const std::vector<B2> &b2, int32_t root) {
  CollapsePlan p;
  p.greedy = getenv("PTC_COLLAPSE") && !strcmp(getenv("PTC_COLLAPSE"), "greedy");
  p.leaf.assign(b2.size(), 0);
  if (p.greedy) return p;
  const double c_t = 1.0, c_n = getenv("PTC_SAH_CN") ? atof(getenv("PTC_SAH_CN")) : 1.8;  // ncu: a node step costs ~1.8 triangle tests per lane
  p.cost.assign(b2.size() * 8, 0.0f);
  p.split.assign(b2.size() * 8, 0);
  // The recurrence only looks down, so the subtrees hanging below binary depth 7 are planned in parallel (each in its own
  // reversed pre-order = children before parents) and the few nodes above them afterwards.
  auto plan_node = [&](const int32_t n) {
    const B2 &nd = b2[n];
    const double area = nd.box.area();
    float *C = &p.cost[(size_t)n * 8];
    uint8_t *S = &p.split[(size_t)n * 8];
    if (nd.left < 0) {
      for (int i = 0; i < 8; i++) C[i] = (float)(area * nd.count * c_t);
      p.leaf[n] = 1;
      return;
    }
    const float *L = &p.cost[(size_t)nd.left * 8], *R = &p.cost[(size_t)nd.right * 8];
    double D[9];
    uint8_t K[9];
    for (int j = 2; j <= 8; j++) {
      D[j] = DBL_MAX, K[j] = 1;
      for (int k = 1; k < j; k++) {
        const double v = (double)L[k - 1] + (double)R[j - k - 1];
        if (v < D[j]) D[j] = v, K[j] = (uint8_t)k;
      }
    }
    const double c_leaf = nd.count <= kLeafMax ? area * nd.count * c_t : DBL_MAX;
    const double c_inner = area * c_n + D[8];
    p.leaf[n] = c_leaf <= c_inner ? 1 : 0;
    C[0] = (float)std::min(c_leaf, c_inner);
    S[7] = K[8];
    for (int i = 2; i <= 7; i++) {
      if (D[i] < (double)C[i - 2]) C[i - 1] = (float)D[i], S[i - 1] = K[i];
      else C[i - 1] = C[i - 2], S[i - 1] = 0;
    }
  };
  auto plan_subtree = [&](const int32_t sub_root) {
    std::vector<int32_t> order, st{sub_root};
    while (!st.empty()) {
      const int32_t n = st.back();
      st.pop_back();
      order.push_back(n);
      if (b2[n].left >= 0) st.push_back(b2[n].left), st.push_back(b2[n].right);
    }
    for (size_t i = order.size(); i-- > 0;) plan_node(order[i]);
  };
  std::vector<int32_t> top, subs;  // nodes above the cut (pre-order), roots of the subtrees below it
  {
    std::vector<std::pair<int32_t, int>> st{{root, 0}};
    while (!st.empty()) {
      const auto [n, d] = st.back();
      st.pop_back();
      if (d >= 7 || b2[n].left < 0) {
        subs.push_back(n);
        continue;
      }
      top.push_back(n);
      st.push_back({b2[n].left, d + 1}), st.push_back({b2[n].right, d + 1});
    }
  }
  {
    const int workers = b2.size() > (1u << 16) ? (int)std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())) : 1;
    std::atomic<size_t> next{0};
    auto run = [&]() {
      for (size_t i = next.fetch_add(1); i < subs.size(); i = next.fetch_add(1)) plan_subtree(subs[i]);
    };
    std::vector<std::thread> th;
    for (int w = 1; w < workers; w++) th.emplace_back(run);
    run();
    for (auto &t : th) t.join();
  }
  for (size_t i = top.size(); i-- > 0;) plan_node(top[i]);
  return p;
}
// the children the plan gives to `slots` slots of a wide node for the subtree of n
void plan_children(const std::vector<B2> &b2, const CollapsePlan &p, int32_t n, int slots, int32_t *ch, int &nch) {
  if (b2[n].left < 0 || slots == 1) {
    ch[nch++] = n;
    return;
  }
  const int k = p.split[(size_t)n * 8 + slots - 1];
  if (k == 0) {
    plan_children(b2, p, n, slots - 1, ch, nch);
    return;
  }
  plan_children(b2, p, b2[n].left, k, ch, nch);
  plan_children(b2, p, b2[n].right, slots - k, ch, nch);
}

void collapse(const std::vector<B2> &b2, int32_t root, const int32_t *prim, const int32_t *live_ids, MeshBuild &m) {
  const auto t_plan = std::chrono::steady_clock::now();
  const CollapsePlan plan = plan_collapse(b2, root);
  if (getenv("PTC_BUILD_TIMING"))
    fprintf(stderr, "[pt_build] collapse plan %.0f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan).count());
  auto is_leaf = [&](int32_t c) { return plan.greedy ? b2[c].left < 0 : plan.leaf[c] != 0; };
  struct Item {
    int32_t b2, wide, depth;
  };
  struct Rec {
    int32_t b2, wide, nch;
    int32_t ch[8];       // children in the order the plan produced them
    int slot_child[8];   // ... and by slot
    uint32_t child_base, tri_base;
  };
  std::vector<Rec> recs;
  m.nodes.clear();
  m.tri48.clear();
  size_t n_nodes = 1, n_tris = 0;
  m.wide_depth = 0;
  // Breadth-first, one wide LEVEL at a time (the tree is ~7 levels deep).  Finding a node's children means chasing the
  // plan through a dozen binary nodes scattered over 100 MB — a DRAM miss each, 60 % of this step when one thread did it
  // for all nodes — and is independent per node: every level does that part in parallel and keeps what the numbering
  // needs (which children are leaves, how many triangles each holds).  The numbering itself — a node's inner children
  // consecutive (provisionally in CHILD order; the slot order the traversal needs comes from the geometry and is fixed up
  // in parallel below), a node's leaf triangles consecutive — is then a sequential pass over compact records: the same
  // numbers as one queue would give.
  struct Found {
    int32_t count[8];
    uint8_t leaf;  // bit i: child i is a leaf
  };
  std::vector<Item> level, next_level;
  std::vector<Found> found;
  level.push_back({root, 0, 0});
  auto find_children = [&](const Item &it, Rec &rec, Found &f) {
    const B2 &n = b2[it.b2];
    int32_t ch[8];
    int nch = 0;
    if (is_leaf(it.b2)) {
      ch[nch++] = it.b2;
    } else if (!plan.greedy) {
      const int k = plan.split[(size_t)it.b2 * 8 + 7];
      plan_children(b2, plan, n.left, k, ch, nch);
      plan_children(b2, plan, n.right, 8 - k, ch, nch);
    } else {
      ch[nch++] = n.left;
      ch[nch++] = n.right;
    }
    while (plan.greedy && nch < 8) {  // open the inner child with the largest surface area
      int pick = -1;
      double pa = -1.0;
      for (int i = 0; i < nch; i++)
        if (b2[ch[i]].left >= 0) {
          const double a = b2[ch[i]].box.area();
          if (a > pa) {
            pa = a;
            pick = i;
          }
        }
      if (pick < 0) break;
      const int32_t p = ch[pick];
      ch[pick] = b2[p].left;
      ch[nch++] = b2[p].right;
    }
    rec.b2 = it.b2, rec.wide = it.wide, rec.nch = nch;
    f.leaf = 0;
    for (int i = 0; i < 8; i++) rec.ch[i] = i < nch ? ch[i] : -1, rec.slot_child[i] = -1, f.count[i] = 0;
    for (int i = 0; i < nch; i++)
      if (is_leaf(ch[i])) f.leaf |= (uint8_t)(1u << i), f.count[i] = b2[ch[i]].count;
  };
  for (int32_t depth = 0; !level.empty(); depth++) {
    m.wide_depth = std::max(m.wide_depth, depth);
    const size_t base = recs.size(), n_level = level.size();
    recs.resize(base + n_level);
    found.resize(n_level);
    const int workers = n_level > 2048 ? (int)std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())) : 1;
    if (workers == 1) {
      for (size_t i = 0; i < n_level; i++) find_children(level[i], recs[base + i], found[i]);
    } else {
      std::vector<std::thread> th;
      const size_t per = (n_level + (size_t)workers - 1) / (size_t)workers;
      for (int w = 0; w < workers; w++) {
        const size_t lo = (size_t)w * per, hi = std::min(n_level, lo + per);
        if (lo < hi)
          th.emplace_back([&, lo, hi]() {
            for (size_t i = lo; i < hi; i++) find_children(level[i], recs[base + i], found[i]);
          });
      }
      for (auto &t : th) t.join();
    }
    next_level.clear();
    for (size_t i = 0; i < n_level; i++) {
      Rec &rec = recs[base + i];
      rec.child_base = (uint32_t)n_nodes, rec.tri_base = (uint32_t)n_tris;
      for (int c = 0; c < rec.nch; c++) {
        if (!(found[i].leaf >> c & 1u)) next_level.push_back({rec.ch[c], (int32_t)n_nodes++, depth + 1});
        else n_tris += (size_t)found[i].count[c];
      }
    }
    level.swap(next_level);
  }
  // recs[k].wide == k for every k (breadth-first numbering): recs[k] describes provisional node k
  std::vector<uint32_t> final_of(n_nodes, 0u);
  auto assign_slots = [&](Rec &rc) {
    const B2 &n = b2[rc.b2];
    const int nch = rc.nch;
    const int32_t *ch = rc.ch;
    // Octant-ordered slots: slot s "lives" at corner (s&1 ? +x : -x, s&2 ? +y : -y, s&4 ? +z : -z); the traversal
    // visits slot (ray octant) first.  Greedy assignment on dot(child centre - node centre, corner direction).
    float nc[3];
    for (int a = 0; a < 3; a++) nc[a] = 0.5f * (n.box.lo[a] + n.box.hi[a]);
    double cost[8][8];
    for (int i = 0; i < nch; i++) {
      const Box3 &cbx = b2[ch[i]].box;
      for (int s = 0; s < 8; s++) {
        double d = 0.0;
        for (int a = 0; a < 3; a++) {
          const double cc = 0.5 * ((double)cbx.lo[a] + cbx.hi[a]) - nc[a];
          d += ((s >> a) & 1) ? cc : -cc;
        }
        cost[i][s] = d;
      }
    }
    int slot_of_child[8];
    bool child_done[8] = {false, false, false, false, false, false, false, false};
    for (int round = 0; round < nch; round++) {
      int bi = -1, bs = -1;
      double bc = -DBL_MAX;
      for (int i = 0; i < nch; i++) {
        if (child_done[i]) continue;
        for (int s = 0; s < 8; s++) {
          if (rc.slot_child[s] >= 0) continue;
          if (cost[i][s] > bc) {
            bc = cost[i][s];
            bi = i;
            bs = s;
          }
        }
      }
      child_done[bi] = true;
      rc.slot_child[bs] = ch[bi];
      slot_of_child[bi] = bs;
    }
    // inner child i holds provisional id child_base + (inner children before it in child order); its final id is
    // child_base + (inner children in lower slots)
    uint32_t prov = rc.child_base;
    for (int i = 0; i < nch; i++) {
      if (is_leaf(ch[i])) continue;
      uint32_t below = 0;
      for (int j = 0; j < nch; j++)
        if (!is_leaf(ch[j]) && slot_of_child[j] < slot_of_child[i]) below++;
      final_of[prov++] = rc.child_base + below;
    }
  };
  auto for_recs = [&](auto &&body) {
    const int workers = recs.size() > 4096 ? (int)std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())) : 1;
    std::vector<std::string> errors((size_t)workers);
    auto run = [&](int w) {
      try {
        for (size_t i = (size_t)w; i < recs.size(); i += (size_t)workers) body(recs[i]);
      } catch (std::exception &e) {
        errors[(size_t)w] = e.what();
      }
    };
    std::vector<std::thread> th;
    for (int w = 1; w < workers; w++) th.emplace_back(run, w);
    run(0);
    for (auto &t : th) t.join();
    for (const std::string &e : errors)
      if (!e.empty()) throw std::runtime_error(e);
  };
  for_recs(assign_slots);
  m.nodes.resize(n_nodes);
  m.tri48.resize(n_tris);
  if (getenv("PTC_BUILD_TIMING"))
    fprintf(stderr, "[pt_build] collapse plan + structure %.0f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan).count());

  auto emit = [&](const Rec &rc) {
    const B2 &n = b2[rc.b2];
    const int *slot_child = rc.slot_child;
    const QFrame fr = make_frame(n.box);
    if (rc.wide == 0)  // the root keeps index 0
      for (int a = 0; a < 3; a++) {
        m.root_lo[a] = fr.origin[a];
        m.root_hi[a] = (float)((double)fr.origin[a] + 255.0 * fr.scale[a]);
        if ((double)m.root_hi[a] < (double)fr.origin[a] + 255.0 * fr.scale[a]) m.root_hi[a] = std::nextafter(m.root_hi[a], INFINITY);
      }
    uint8_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t qlo[3][8], qhi[3][8];
    for (int a = 0; a < 3; a++)
      for (int s = 0; s < 8; s++) {
        qlo[a][s] = 255;
        qhi[a][s] = 0;
      }
    uint32_t imask = 0;
    const uint32_t child_base = rc.child_base, tri_base = rc.tri_base;
    int tri_off = 0;
    for (int s = 0; s < 8; s++) {
      const int32_t c = slot_child[s];
      if (c < 0) continue;
      const B2 &cn = b2[c];
      if (!is_leaf(c)) {
        imask |= 1u << s;
        meta[s] = (uint8_t)(0x20 | (24 + s));
      } else {
        const int cnt = cn.count;
        const uint32_t unary = cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u);
        meta[s] = (uint8_t)((unary << 5) | (uint32_t)tri_off);
        for (int k = 0; k < cnt; k++) {
          const int32_t orig = live_ids[prim[cn.first + k]];
          const float *t = tri_ptr(m, orig);
          Tri48 r;
          // edge vectors with the reference's own subtraction (bvh.rs:94-95), so precomputing them changes no bit
          r.t[0] = make_float4(t[0], t[1], t[2], u2f_host((uint32_t)orig));
          r.t[1] = make_float4(t[3] - t[0], t[4] - t[1], t[5] - t[2], u2f_host((uint32_t)m.order[orig]));
          r.t[2] = make_float4(t[6] - t[0], t[7] - t[1], t[8] - t[2], 0.0f);
          m.tri48[(size_t)tri_base + (size_t)tri_off + (size_t)k] = r;
        }
        tri_off += cnt;
      }
      for (int a = 0; a < 3; a++) {
        const float clo = cn.box.lo[a], chi = cn.box.hi[a];
        const double padc = (double)std::max(std::fabs(clo), std::fabs(chi)) * 0x1p-21 + 1e-30;
        double ql = std::floor(((double)clo - padc - (double)fr.origin[a]) / fr.scale[a]);
        double qh = std::ceil(((double)chi + padc - (double)fr.origin[a]) / fr.scale[a]);
        ql = std::min(std::max(ql, 0.0), 255.0);
        qh = std::min(std::max(qh, 0.0), 255.0);
        qlo[a][s] = (uint8_t)ql;
        qhi[a][s] = (uint8_t)qh;
      }
    }
    auto pack4 = [](const uint8_t *b) -> float {
      return u2f_host((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24));
    };
    Node8 out;
    const uint32_t ebits = (uint32_t)fr.biased_exp[0] | ((uint32_t)fr.biased_exp[1] << 8) |
                           ((uint32_t)fr.biased_exp[2] << 16) | (imask << 24);
    out.q[0] = make_float4(fr.origin[0], fr.origin[1], fr.origin[2], u2f_host(ebits));
    out.q[1] = make_float4(u2f_host(child_base), u2f_host(tri_base), pack4(meta), pack4(meta + 4));
    out.q[2] = make_float4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
    out.q[3] = make_float4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    out.q[4] = make_float4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    m.nodes[(size_t)final_of[(size_t)rc.wide]] = out;
  };
  // frames, quantisation and triangle records: independent per wide node
  for_recs(emit);
  (void)f2u_host;
}

}  // namespace

void build_mesh(MeshBuild &m, int threads) {
  const int64_t n = m.n;
  if (n <= 0) throw std::runtime_error("mesh without triangles");
  if (n > (int64_t)0x7fffffff) throw std::runtime_error("mesh too large");
  int nt = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
  int par_levels = 0;
  while ((1 << par_levels) < nt && par_levels < 6) par_levels++;

  // PTC_BUILD_TIMING=1: per-step wall time on stderr (DESIGN.md section 5 quotes it)
  const bool timing = getenv("PTC_BUILD_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(now() - t0).count(); };
  auto t_begin = now();

  // ---- step 1
  std::vector<float> cen((size_t)n * 3);
  for (int64_t i = 0; i < n; i++) {
    const float *t = tri_ptr(m, i);
    for (int a = 0; a < 3; a++) cen[(size_t)i * 3 + a] = ((t[a] + t[3 + a]) + t[6 + a]) * (1.0f / 3.0f);  // bvh.rs:46
  }
  if (!m.ref_done) {
    m.dead.assign((size_t)n, 0);
    m.order.assign((size_t)n, 0);
  }
  if (!m.ref_done) {
    std::vector<int64_t> idx((size_t)n);
    for (int64_t i = 0; i < n; i++) idx[(size_t)i] = i;
    RefCtx rc;
    rc.m = &m;
    rc.cen = cen.data();
    rc.dead = m.dead.data();
    const TieMode tm = tie_mode_from_env();
    rc.tie_kind = tm.kind, rc.tie_seed = tm.seed;
    std::vector<uint64_t> scratch(n >= 4096 ? (size_t)n * 2 : 0);
    ref_build(rc, idx.data(), n, 0, false, par_levels, scratch.empty() ? nullptr : scratch.data());
    m.ref_nodes = rc.nodes.load();
    m.ref_leaves = rc.leaves.load();
    m.ref_depth = rc.max_depth.load();
    for (int64_t pos = 0; pos < n; pos++) m.order[(size_t)idx[(size_t)pos]] = (int32_t)pos;
  }

  const double ms_ref = ms_since(t_begin);
  auto t_sah = now();

  // ---- normals by position in the reference's DFS leaf order (the traversal's result key carries that position), the
  // original triangle index in .w
  m.normals.resize((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    const float *t = tri_ptr(m, i);
    m.normals[(size_t)m.order[(size_t)i]] = make_float4(t[9], t[10], t[11], u2f_host((uint32_t)i));
  }

  // ---- step 2
  std::vector<int32_t> live_ids;
  live_ids.reserve((size_t)n);
  for (int64_t i = 0; i < n; i++)
    if (!m.dead[(size_t)i]) live_ids.push_back((int32_t)i);
  m.live = (int64_t)live_ids.size();
  m.nodes.clear();
  m.tri48.clear();
  m.wide_depth = 0;
  if (live_ids.empty()) {
    // every triangle is unreachable in the reference: an empty root (no child bits) never reports a hit
    Node8 empty;
    const uint32_t e = 127u | (127u << 8) | (127u << 16);
    empty.q[0] = make_float4(0, 0, 0, u2f_host(e));
    empty.q[1] = make_float4(u2f_host(1u), u2f_host(0u), u2f_host(0u), u2f_host(0u));
    const float inv = u2f_host(0xffffffffu), zero = u2f_host(0u);
    empty.q[2] = make_float4(inv, inv, inv, inv);
    empty.q[3] = make_float4(inv, inv, zero, zero);
    empty.q[4] = make_float4(zero, zero, zero, zero);
    m.nodes.push_back(empty);
    Tri48 z;
    z.t[0] = z.t[1] = z.t[2] = make_float4(0, 0, 0, 0);
    m.tri48.push_back(z);  // keep the buffer non-empty
    for (int a = 0; a < 3; a++) m.root_lo[a] = 1.0f, m.root_hi[a] = -1.0f;
    m.built = true;
    return;
  }
  const size_t nl = live_ids.size();
  std::vector<Box3> pbox(nl);
  std::vector<float> pcen(nl * 3);
  std::vector<int32_t> prim(nl);
  for (size_t i = 0; i < nl; i++) {
    const float *t = tri_ptr(m, live_ids[i]);
    pbox[i].reset();
    pbox[i].grow(t);
    pbox[i].grow(t + 3);
    pbox[i].grow(t + 6);
    for (int a = 0; a < 3; a++) pcen[i * 3 + a] = 0.5f * (pbox[i].lo[a] + pbox[i].hi[a]);
    prim[i] = (int32_t)i;
  }
  SahCtx sc;
  sc.nodes.resize(2 * nl);
  sc.pbox = pbox.data();
  sc.pcen = pcen.data();
  sc.prim = prim.data();
  const int32_t root = sah_build(sc, 0, (int32_t)nl, par_levels);
  const double ms_sah = ms_since(t_sah);
  auto t_col = now();

  // ---- step 3
  collapse(sc.nodes, root, prim.data(), live_ids.data(), m);
  m.built = true;
  if (timing) {
    size_t inner = 0, leaf = 0, tris = 0;
    for (const Node8 &nd : m.nodes) {
      uint32_t w[2];
      memcpy(w, &nd.q[1].z, 8);
      for (int s8 = 0; s8 < 8; s8++) {
        const uint32_t meta = (w[s8 >> 2] >> (8 * (s8 & 3))) & 0xffu;
        if (meta == 0) continue;
        if ((meta >> 5) == 1u && (meta & 0x1fu) >= 24u) inner++;  // inner child: unary count 001, slot index + 24 (pt_bvh8.h)
        else leaf++, tris += (meta >> 5) == 1 ? 1 : ((meta >> 5) == 3 ? 2 : 3);
      }
    }
    fprintf(stderr, "[pt_build] wide nodes %zu: %.2f children per node (%.2f inner + %.2f leaves), %.2f triangles per leaf, depth %d\n",
            m.nodes.size(), (double)(inner + leaf) / m.nodes.size(), (double)inner / m.nodes.size(), (double)leaf / m.nodes.size(),
            (double)tris / std::max<size_t>(1, leaf), m.wide_depth);
  }
  if (timing)
    fprintf(stderr, "[pt_build] %lld triangles (%lld live), %d threads: reference-BVH restatement %.0f ms, binned SAH %.0f ms, "
                    "8-wide collapse + quantisation %.0f ms; %zu wide nodes\n",
            (long long)n, (long long)m.live, nt, ms_ref, ms_sah, ms_since(t_col), m.nodes.size());
}

}  // namespace pt
