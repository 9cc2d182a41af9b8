// pt_build_dev.cu — see pt_build_dev.h.  CUDA for sm_100a; CUB (header-only, part of the toolkit) provides the radix sort.
//
// Replaces, for large meshes, the host recursion of pt_build.cpp — itself the restatement of the reference's host-side
// recursive sort, BVHNode::new (/root/reference/src/acceleration/bvh.rs:15-76, sort at :45-53).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "pt_build_dev.h"

namespace pt {
namespace {

#define CKB(call)                                                                                                        \
  do {                                                                                                                   \
    cudaError_t e_ = (call);                                                                                             \
    if (e_ != cudaSuccess)                                                                                               \
      throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (pt_build_dev.cu:" + std::to_string(__LINE__) + ")"); \
  } while (0)

template <typename T>
struct Buf {
  T *p = nullptr;
  size_t n = 0;
  explicit Buf(size_t count = 0) {
    if (count) alloc(count);
  }
  void alloc(size_t count) {
    release();
    CKB(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    n = count;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
  }
  T *take() {
    T *r = p;
    p = nullptr;
    return r;
  }
  ~Buf() { release(); }
  Buf(const Buf &) = delete;
  Buf &operator=(const Buf &) = delete;
};

// Temporaries come out of ONE allocation per build step: cudaMalloc / cudaFree synchronise the device and their cost
// varies by an order of magnitude from call to call (measured: 137 to 495 ms for the same 2 M-triangle commit with ~25
// separate buffers), which is the same order as the whole build.
struct Arena {
  char *base = nullptr;
  size_t size = 0, used = 0;
  explicit Arena(size_t bytes) : size(bytes) { CKB(cudaMalloc(&base, std::max<size_t>(bytes, 256))); }
  ~Arena() {
    if (base) cudaFree(base);
  }
  Arena(const Arena &) = delete;
  Arena &operator=(const Arena &) = delete;
  static size_t need(size_t count, size_t elem) { return (count * elem + 255) & ~(size_t)255; }
  template <typename T>
  T *get(size_t count) {
    const size_t b = need(std::max<size_t>(count, 1), sizeof(T));
    if (used + b > size) throw std::runtime_error("pt_build_dev: arena too small");
    T *p = reinterpret_cast<T *>(base + used);
    used += b;
    return p;
  }
};

constexpr int kThreads = 256;
inline unsigned blocks_for(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

// order-preserving integer image of a float; -0.0 keyed as +0.0 when `canon` (the sort's comparator calls them equal)
__device__ __forceinline__ uint32_t fkey(float f, bool canon) {
  if (canon && f == 0.0f) f = 0.0f;
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// The node of depth L that holds position p.  The shape of the reference tree is a function of the triangle count alone:
// a node of size s > 4 (and depth < 25) splits at s / 2 (bvh.rs:19-27,55-56).  `real` is false when an ancestor was
// already a leaf: the descent then continues through VIRTUAL nodes, which only serve to give every position a sort key
// that keeps it where it is.
struct Seg {
  uint32_t lo, hi, idx;
  bool real;
};
__device__ __forceinline__ Seg descend(uint32_t p, uint32_t n, int L) {
  Seg s{0u, n, 0u, true};
  for (int d = 0; d < L; d++) {
    const uint32_t size = s.hi - s.lo;
    if (size <= 4u || d >= 25) s.real = false;
    const uint32_t mid = s.lo + size / 2u;
    if (p < mid) s.hi = mid, s.idx = s.idx * 2u;
    else s.lo = mid, s.idx = s.idx * 2u + 1u;
  }
  return s;
}

// per triangle: centroid ((v0 + v1) + v2) * (1/3) (bvh.rs:46; this TU is compiled with --fmad=false) and vertex bounds
__global__ void k_tri_prepare(const float *tris, uint32_t n, float *cen, float *tbox, uint32_t *idx, uint8_t *dead) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *t = tris + (size_t)i * 12;
  for (int a = 0; a < 3; a++) {
    cen[(size_t)i * 3 + a] = ((t[a] + t[3 + a]) + t[6 + a]) * (1.0f / 3.0f);
    tbox[(size_t)i * 6 + a] = fminf(fminf(t[a], t[3 + a]), t[6 + a]);
    tbox[(size_t)i * 6 + 3 + a] = fmaxf(fmaxf(t[a], t[3 + a]), t[6 + a]);
  }
  idx[i] = i;
  dead[i] = 0;
}

__global__ void k_fill_u32(uint32_t *p, size_t n, uint32_t lo_value, uint32_t hi_value) {  // segment bounds: [lo x3, hi x3] per node
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (i % 6) < 3 ? lo_value : hi_value;
}

// vertex bounds of every real node of depth L (bvh.rs:17): warp-aggregated atomic min / max on ordered integer keys
__global__ void k_level_bounds(const uint32_t *idx, const float *tbox, uint32_t n, int L, uint32_t *segb) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  Seg s{0u, 0u, 0u, false};
  if (p < n) s = descend(p, n, L);
  const bool part = p < n && s.real;
  const uint32_t mask = __ballot_sync(0xffffffffu, part);
  if (!part) return;
  const float *b = tbox + (size_t)idx[p] * 6;
  uint32_t k[6];
  for (int a = 0; a < 6; a++) k[a] = fkey(b[a], false);
  const int leader = __ffs((int)mask) - 1;
  const uint32_t lead_idx = __shfl_sync(mask, s.idx, leader);
  uint32_t *dst = segb + (size_t)s.idx * 6;
  if (__all_sync(mask, s.idx == lead_idx)) {
    for (int a = 0; a < 3; a++) k[a] = __reduce_min_sync(mask, k[a]);
    for (int a = 3; a < 6; a++) k[a] = __reduce_max_sync(mask, k[a]);
    if ((int)(threadIdx.x & 31u) == leader) {
      for (int a = 0; a < 3; a++) atomicMin(dst + a, k[a]);
      for (int a = 3; a < 6; a++) atomicMax(dst + a, k[a]);
    }
  } else {
    for (int a = 0; a < 3; a++) atomicMin(dst + a, k[a]);
    for (int a = 3; a < 6; a++) atomicMax(dst + a, k[a]);
  }
}

// flat test (aabb.rs:40: a zero-extent node is never entered), split axis (bvh.rs:36-43) and the level's sort key
__global__ void k_level_keys(const uint32_t *idx, const float *cen, const uint32_t *segb, uint32_t n, int L, uint8_t *dead,
                             unsigned long long *keys) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const Seg s = descend(p, n, L);
  const uint32_t tri = idx[p];
  uint32_t key = p - s.lo;  // leaves and virtual nodes: stay in place
  if (s.real) {
    const uint32_t *b = segb + (size_t)s.idx * 6;
    const float lx = fkey_inv(b[0]), ly = fkey_inv(b[1]), lz = fkey_inv(b[2]), hx = fkey_inv(b[3]), hy = fkey_inv(b[4]), hz = fkey_inv(b[5]);
    if (lx == hx || ly == hy || lz == hz) dead[tri] = 1;
    const uint32_t size = s.hi - s.lo;
    if (size > 4u && L < 25 && keys != nullptr) {
      const float ex = hx - lx, ey = hy - ly, ez = hz - lz;
      const int axis = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);
      key = fkey(cen[(size_t)tri * 3 + axis], true);
    }
  }
  if (keys) keys[p] = ((unsigned long long)s.idx << 32) | (unsigned long long)key;
}

__global__ void k_finish_order(const uint32_t *idx, const uint8_t *dead, uint32_t n, int32_t *order, uint8_t *dead_by_pos) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t tri = idx[p];
  order[tri] = (int32_t)p;
  dead_by_pos[p] = dead[tri];
}

// normals by DFS position, original index in .w (pt_build.h)
__global__ void k_normals(const float *tris, const int32_t *order, uint32_t n, float4 *normals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *t = tris + (size_t)i * 12;
  normals[order[i]] = make_float4(t[9], t[10], t[11], __uint_as_float(i));
}

// ---- step 2: the wide tree ------------------------------------------------------------------------------------
// Structure of one wide node, computed on the host from the live counts alone (gen_structure below).
struct WNode {
  uint32_t child_lo[8];   // first live rank of the child's range (its triangles are live_tri[child_lo .. child_lo + cnt))
  uint32_t child_cnt[8];  // live triangles: 0 = no child, 1..3 = leaf, > 3 = inner node
  uint32_t child_base;    // index of the first inner child (inner children are consecutive, in child order)
  uint32_t tri_base;      // first Tri48 record of this node's leaf children
};

struct Box {
  float lo[3], hi[3];
};

__device__ __forceinline__ void box_grow(Box &b, const float *p) {
  for (int a = 0; a < 3; a++) b.lo[a] = fminf(b.lo[a], p[a]), b.hi[a] = fmaxf(b.hi[a], p[a]);
}

// A reference leaf with four live triangles becomes two leaves of two (a wide-node leaf holds at most three).  WHICH two go
// together is free — the records carry their own ids — so each quad is re-paired to the split with the smallest summed box
// area: on a regular mesh the reference's centroid order pairs triangles of neighbouring cells as often as the two halves
// of one cell, and a mismatched pair doubles the leaf's box.
__device__ __forceinline__ float pair_area(const float *a, const float *b) {
  float lo[3], hi[3];
  for (int k = 0; k < 3; k++) {
    lo[k] = fminf(fminf(fminf(a[k], a[3 + k]), a[6 + k]), fminf(fminf(b[k], b[3 + k]), b[6 + k]));
    hi[k] = fmaxf(fmaxf(fmaxf(a[k], a[3 + k]), a[6 + k]), fmaxf(fmaxf(b[k], b[3 + k]), b[6 + k]));
  }
  const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
  return x * y + y * z + z * x;
}
__global__ void k_pair_quads(const uint32_t *quads, uint32_t nq, uint32_t *live_tri, const float *tris) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint32_t *t = live_tri + quads[q];
  const uint32_t i0 = t[0], i1 = t[1], i2 = t[2], i3 = t[3];
  const float *p0 = tris + (size_t)i0 * 12, *p1 = tris + (size_t)i1 * 12, *p2 = tris + (size_t)i2 * 12, *p3 = tris + (size_t)i3 * 12;
  const float a01 = pair_area(p0, p1) + pair_area(p2, p3), a02 = pair_area(p0, p2) + pair_area(p1, p3), a03 = pair_area(p0, p3) + pair_area(p1, p2);
  if (a02 < a01 && a02 <= a03) t[1] = i2, t[2] = i1;       // (0,2) (1,3)
  else if (a03 < a01 && a03 < a02) t[1] = i3, t[3] = i1;   // (0,3) (2,1)
}

// Bottom-up, one thread per wide node of a level (deepest level first), indexed by STRUCTURAL id (the breadth-first
// numbering of gen_structure): the node's box = union of its children's boxes (from triangles or from the level below).
__global__ void k_boxes_level(const WNode *wn, uint32_t first, uint32_t count, const uint32_t *live_tri, const float *tris, Box *nodebox) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const uint32_t node = first + t;
  const WNode w = wn[node];
  Box nb;
  for (int a = 0; a < 3; a++) nb.lo[a] = INFINITY, nb.hi[a] = -INFINITY;
  uint32_t inner_rank = 0;
  for (int c = 0; c < 8; c++) {
    const uint32_t cnt = w.child_cnt[c];
    if (cnt == 0u) break;  // children are packed at the front
    if (cnt > 3u) {
      const Box b = nodebox[w.child_base + inner_rank++];
      for (int a = 0; a < 3; a++) nb.lo[a] = fminf(nb.lo[a], b.lo[a]), nb.hi[a] = fmaxf(nb.hi[a], b.hi[a]);
    } else {
      for (uint32_t k = 0; k < cnt; k++) {
        const float *tp = tris + (size_t)live_tri[w.child_lo[c] + k] * 12;
        box_grow(nb, tp), box_grow(nb, tp + 3), box_grow(nb, tp + 6);
      }
    }
  }
  nodebox[node] = nb;
}

// inner children of the structural node that sits at final position f of its level
__global__ void k_inner_count(const WNode *wn, const uint32_t *sid, uint32_t count, uint32_t *out) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= count) return;
  const WNode &w = wn[sid[f]];
  uint32_t k = 0;
  for (int c = 0; c < 8; c++) k += w.child_cnt[c] > 3u ? 1u : 0u;
  out[f] = k;
}

// Top-down, one thread per wide node of a level.  The traversal finds the inner child in slot s at
// child_base + popc(inner slots below s) (pt_bvh8.h), so a node's inner children must be numbered in SLOT order, and the
// slots come from the geometry (octant of the child's centre).  `sid[f]` = structural id of the node whose FINAL index is
// level_first + f; this kernel assigns the slots, writes the Node8 at the final index and hands its inner children, in
// slot order, the final positions cbase[f] + k of the next level (`sid_next`).  Also writes the Tri48 records of the leaf
// children (their place, tri_base, is structural: any place is as good as another).
__global__ void k_emit_level(const WNode *wn, const uint32_t *sid, uint32_t level_first, uint32_t count, const uint32_t *cbase,
                             uint32_t next_first, uint32_t *sid_next, const uint32_t *live_tri, const float *tris, const int32_t *order,
                             const Box *nodebox, float4 *nodes, float4 *tri48, float *root_frame, int *error) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= count) return;
  const uint32_t u = sid[f];
  const WNode w = wn[u];
  const Box nb = nodebox[u];
  Box cb[8];
  uint32_t child_sid[8];
  bool inner[8];
  int nch = 0;
  uint32_t inner_rank = 0;
  for (int c = 0; c < 8; c++) {
    const uint32_t cnt = w.child_cnt[c];
    if (cnt == 0u) break;
    nch = c + 1;
    inner[c] = cnt > 3u;
    child_sid[c] = 0u;
    if (inner[c]) {
      child_sid[c] = w.child_base + inner_rank++;
      cb[c] = nodebox[child_sid[c]];
    } else {
      for (int a = 0; a < 3; a++) cb[c].lo[a] = INFINITY, cb[c].hi[a] = -INFINITY;
      for (uint32_t k = 0; k < cnt; k++) {
        const float *tp = tris + (size_t)live_tri[w.child_lo[c] + k] * 12;
        box_grow(cb[c], tp), box_grow(cb[c], tp + 3), box_grow(cb[c], tp + 6);
      }
    }
  }

  // octant-ordered slots: slot s lives at corner (s&1 ? +x : -x, s&2 ? +y : -y, s&4 ? +z : -z); greedy on the projection
  // of (child centre - node centre) on the corner direction (same rule as pt_build.cpp: collapse)
  int slot_child[8];
  for (int s = 0; s < 8; s++) slot_child[s] = -1;
  {
    double ccx[8], ccy[8], ccz[8];
    const double ncx = 0.5f * (nb.lo[0] + nb.hi[0]), ncy = 0.5f * (nb.lo[1] + nb.hi[1]), ncz = 0.5f * (nb.lo[2] + nb.hi[2]);
    for (int i = 0; i < nch; i++) {
      ccx[i] = 0.5 * ((double)cb[i].lo[0] + cb[i].hi[0]) - ncx;
      ccy[i] = 0.5 * ((double)cb[i].lo[1] + cb[i].hi[1]) - ncy;
      ccz[i] = 0.5 * ((double)cb[i].lo[2] + cb[i].hi[2]) - ncz;
    }
    uint32_t child_done = 0u;
    for (int round = 0; round < nch; round++) {
      int bi = -1, bs = -1;
      double bc = -1.7976931348623157e308;
      for (int i = 0; i < nch; i++) {
        if (child_done & (1u << i)) continue;
        for (int s = 0; s < 8; s++) {
          if (slot_child[s] >= 0) continue;
          const double d = ((s & 1) ? ccx[i] : -ccx[i]) + ((s & 2) ? ccy[i] : -ccy[i]) + ((s & 4) ? ccz[i] : -ccz[i]);
          if (d > bc) bc = d, bi = i, bs = s;
        }
      }
      child_done |= 1u << bi;
      slot_child[bs] = bi;
    }
  }

  // frame = node box padded on every side (pt_build.cpp: make_frame)
  float origin[3];
  int biased[3];
  double scale[3];
  for (int a = 0; a < 3; a++) {
    const float lo = nb.lo[a], hi = nb.hi[a];
    const float ext = hi - lo;
    const float maxabs = fmaxf(fabsf(lo), fabsf(hi));
    const float pad = ext * (1.0f / 512.0f) + maxabs * 0x1p-20f + 1e-30f;
    float flo = lo - pad, fhi = hi + pad;
    if (!(flo < lo)) flo = nextafterf(lo, -INFINITY);
    if (!(fhi > hi)) fhi = nextafterf(hi, INFINITY);
    const double need = ((double)fhi - (double)flo) / 255.0;
    int ex;
    frexp(need, &ex);
    int b = ex + 127;
    if (b < 1) b = 1;
    if (b > 254 - 15) {
      *error = 1;  // mesh extent too large for the quantised frame (trav_node adds 15 to the exponent)
      b = 254 - 15;
    }
    origin[a] = flo;
    biased[a] = b;
    scale[a] = ldexp(1.0, b - 127);
  }
  if (level_first == 0u && f == 0u) {
    for (int a = 0; a < 3; a++) {
      root_frame[a] = origin[a];
      const double top = (double)origin[a] + 255.0 * scale[a];
      float h = (float)top;
      if ((double)h < top) h = nextafterf(h, INFINITY);
      root_frame[3 + a] = h;
    }
  }

  uint32_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0}, qlo[3][8], qhi[3][8];
  for (int a = 0; a < 3; a++)
    for (int s = 0; s < 8; s++) qlo[a][s] = 255u, qhi[a][s] = 0u;
  uint32_t imask = 0u, tri_off = 0u, inner_seen = 0u;
  const uint32_t my_cbase = cbase[f];
  for (int s = 0; s < 8; s++) {
    const int c = slot_child[s];
    if (c < 0) continue;
    if (inner[c]) {
      imask |= 1u << s;
      meta[s] = 0x20u | (24u + (uint32_t)s);
      sid_next[my_cbase + inner_seen++] = child_sid[c];
    } else {
      const uint32_t cnt = w.child_cnt[c];
      const uint32_t unary = cnt == 1u ? 1u : (cnt == 2u ? 3u : 7u);
      meta[s] = (unary << 5) | tri_off;
      for (uint32_t k = 0; k < cnt; k++) {
        const uint32_t orig = live_tri[w.child_lo[c] + k];
        const float *tp = tris + (size_t)orig * 12;
        float4 *r = tri48 + (size_t)(w.tri_base + tri_off + k) * 3;
        // edge vectors with the reference's own subtraction (bvh.rs:94-95), so precomputing them changes no bit
        r[0] = make_float4(tp[0], tp[1], tp[2], __uint_as_float(orig));
        r[1] = make_float4(tp[3] - tp[0], tp[4] - tp[1], tp[5] - tp[2], __uint_as_float((uint32_t)order[orig]));
        r[2] = make_float4(tp[6] - tp[0], tp[7] - tp[1], tp[8] - tp[2], 0.0f);
      }
      tri_off += cnt;
    }
    for (int a = 0; a < 3; a++) {
      const float clo = cb[c].lo[a], chi = cb[c].hi[a];
      const double padc = (double)fmaxf(fabsf(clo), fabsf(chi)) * 0x1p-21 + 1e-30;
      double ql = floor(((double)clo - padc - (double)origin[a]) / scale[a]);
      double qh = ceil(((double)chi + padc - (double)origin[a]) / scale[a]);
      ql = fmin(fmax(ql, 0.0), 255.0);
      qh = fmin(fmax(qh, 0.0), 255.0);
      qlo[a][s] = (uint32_t)ql;
      qhi[a][s] = (uint32_t)qh;
    }
  }

  auto pack4 = [](const uint32_t *b) { return __uint_as_float(b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24)); };
  const uint32_t ebits = (uint32_t)biased[0] | ((uint32_t)biased[1] << 8) | ((uint32_t)biased[2] << 16) | (imask << 24);
  float4 *out = nodes + (size_t)(level_first + f) * 5;
  out[0] = make_float4(origin[0], origin[1], origin[2], __uint_as_float(ebits));
  out[1] = make_float4(__uint_as_float(next_first + my_cbase), __uint_as_float(w.tri_base), pack4(meta), pack4(meta + 4));
  out[2] = make_float4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
  out[3] = make_float4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
  out[4] = make_float4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
}

// ---- host side ----------------------------------------------------------------------------------------------------
// The wide tree's structure from the live counts alone.  Binary tree = the reference's own (split at size / 2 while
// size > 4, bvh.rs:19-27,55-56) with every range measured in LIVE triangles (R = live rank of a DFS position); ranges
// with <= 3 live triangles are leaves, a reference leaf that still holds 4 live ones is halved once more.  A wide node
// takes up to three binary levels; a split with an empty side costs no level.  Breadth-first numbering: a node's inner
// children get consecutive ids, every level is a contiguous id range.
struct Item {
  uint32_t ref_lo, ref_hi, live_lo, live_hi;
  int32_t depth;  // reference depth; -1 = below a reference leaf (split by live count)
};
struct Structure {
  std::vector<uint32_t> quads;  // first live rank of every reference leaf that holds 4 live triangles (see k_pair_quads)
  std::vector<WNode> nodes;
  std::vector<uint32_t> level_first;  // level l = ids [level_first[l], level_first[l + 1])
  uint32_t n_tri_records = 0;
};

void expand(const Item &it, int levels, const std::vector<uint32_t> &R, std::vector<Item> &out, std::vector<uint32_t> &quads) {
  const uint32_t live = it.live_hi - it.live_lo;
  if (live == 0u) return;
  if (live <= 3u || levels == 0) {
    out.push_back(it);
    return;
  }
  Item l = it, r = it;
  const uint32_t rsize = it.ref_hi - it.ref_lo;
  if (it.depth >= 0 && rsize > 4u && it.depth < 25) {
    const uint32_t mid = it.ref_lo + rsize / 2u;
    l.ref_hi = mid, r.ref_lo = mid;
    l.depth = r.depth = it.depth + 1;
    l.live_hi = r.live_lo = R[mid];
  } else {
    if (it.depth >= 0 && live == 4u) quads.push_back(it.live_lo);
    const uint32_t lm = it.live_lo + live / 2u;
    l.depth = r.depth = -1;
    l.live_hi = r.live_lo = lm;
  }
  if (l.live_hi == l.live_lo) expand(r, levels, R, out, quads);
  else if (r.live_hi == r.live_lo) expand(l, levels, R, out, quads);
  else expand(l, levels - 1, R, out, quads), expand(r, levels - 1, R, out, quads);
}

Structure gen_structure(uint32_t n, const std::vector<uint32_t> &R) {
  Structure st;
  std::vector<Item> queue{Item{0u, n, 0u, R[n], 0}};
  st.level_first.push_back(0u);
  size_t level_end = 1;
  std::vector<Item> ch;
  for (size_t qi = 0; qi < queue.size(); qi++) {
    if (qi == level_end) {
      st.level_first.push_back((uint32_t)qi);
      level_end = queue.size();
    }
    const Item it = queue[qi];
    ch.clear();
    if (it.live_hi - it.live_lo <= 3u) ch.push_back(it);  // a mesh of <= 3 live triangles: a root with one leaf
    else expand(it, 3, R, ch, st.quads);
    WNode w;
    memset(&w, 0, sizeof(w));
    w.child_base = (uint32_t)queue.size();
    w.tri_base = st.n_tri_records;
    for (size_t c = 0; c < ch.size(); c++) {
      const uint32_t live = ch[c].live_hi - ch[c].live_lo;
      w.child_lo[c] = ch[c].live_lo;
      w.child_cnt[c] = live;
      if (live > 3u) queue.push_back(ch[c]);
      else st.n_tri_records += live;
    }
    st.nodes.push_back(w);
  }
  st.level_first.push_back((uint32_t)queue.size());
  return st;
}

// nodes / leaves / depth of the reference tree (they depend on n only)
void ref_counts(uint64_t n, int depth, int64_t &nodes, int64_t &leaves, int32_t &max_depth, std::unordered_map<uint64_t, std::pair<int64_t, int64_t>> &memo) {
  max_depth = std::max(max_depth, depth);
  if (n <= 4 || depth >= 25) {
    nodes += 1, leaves += 1;
    return;
  }
  if (depth < 20) {  // memoised by size where the depth cap cannot interfere (n < 2^27)
    auto f = memo.find(n);
    if (f != memo.end()) {
      nodes += f->second.first, leaves += f->second.second;
      int32_t d = depth;
      for (uint64_t s = n; s > 4; s = s - s / 2) d++;
      max_depth = std::max(max_depth, d);
      return;
    }
  }
  int64_t a = 1, b = 0;
  ref_counts(n / 2, depth + 1, a, b, max_depth, memo);
  ref_counts(n - n / 2, depth + 1, a, b, max_depth, memo);
  if (depth < 20) memo[n] = {a, b};
  nodes += a, leaves += b;
}

double ms_between(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
  return std::chrono::duration<double, std::milli>(b - a).count();
}

}  // namespace

void build_mesh_device(MeshBuild &m, DevMeshBuffers &out, DevBuildTiming *timing, bool ref_only) {
  const int64_t n64 = m.n;
  if (n64 <= 0) throw std::runtime_error("mesh without triangles");
  if (n64 >= ((int64_t)1 << 27)) throw std::runtime_error("mesh too large for the device builder");
  const uint32_t n = (uint32_t)n64;
  auto now = [] { return std::chrono::steady_clock::now(); };
  const auto t0 = now();
  int device = 0;
  CKB(cudaGetDevice(&device));
  cudaStream_t stream = nullptr;  // legacy default stream: the build is a synchronous host call

  // ---- upload
  int sort_levels = 0;  // levels of the reference tree that still have a node to split
  while (sort_levels < 25 && (((uint64_t)n + ((1ull << sort_levels) - 1)) >> sort_levels) > 4ull) sort_levels++;
  size_t temp_bytes = 0;
  {
    cub::DoubleBuffer<unsigned long long> kb(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> vb(nullptr, nullptr);
    CKB(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, kb, vb, (int)n, 0, 64, stream));
  }
  const size_t segb_count = ((size_t)1 << sort_levels) * 6;
  Arena a1(Arena::need((size_t)n * 12, 4) + Arena::need((size_t)n * 3, 4) + Arena::need((size_t)n * 6, 4) + 2 * Arena::need(n, 4) +
           2 * Arena::need(n, 8) + 2 * Arena::need(n, 1) + Arena::need(n, 4) + Arena::need(segb_count, 4) + Arena::need(temp_bytes, 1) + 4096);
  struct P {
    float *p;
  } d_tris{a1.get<float>((size_t)n * 12)}, d_cen{a1.get<float>((size_t)n * 3)}, d_tbox{a1.get<float>((size_t)n * 6)};
  struct PU {
    uint32_t *p;
  } d_idx[2] = {{a1.get<uint32_t>(n)}, {a1.get<uint32_t>(n)}}, d_segb{a1.get<uint32_t>(segb_count)};
  struct PK {
    unsigned long long *p;
  } d_keys[2] = {{a1.get<unsigned long long>(n)}, {a1.get<unsigned long long>(n)}};
  struct PB {
    uint8_t *p;
  } d_dead{a1.get<uint8_t>(n)}, d_dead_pos{a1.get<uint8_t>(n)}, d_temp{a1.get<uint8_t>(temp_bytes)};
  struct PI {
    int32_t *p;
  } d_order{a1.get<int32_t>(n)};
  CKB(cudaMemcpyAsync(d_tris.p, m.tris.data(), (size_t)n * 48, cudaMemcpyHostToDevice, stream));
  k_tri_prepare<<<blocks_for(n), kThreads, 0, stream>>>(d_tris.p, n, d_cen.p, d_tbox.p, d_idx[0].p, d_dead.p);
  CKB(cudaGetLastError());
  CKB(cudaStreamSynchronize(stream));
  const auto t1 = now();

  // ---- step 1: the reference tree, one level per pass
  int cur = 0;
  for (int L = 0; L <= sort_levels; L++) {
    const size_t segs = (size_t)1 << L;
    k_fill_u32<<<blocks_for(segs * 6), kThreads, 0, stream>>>(d_segb.p, segs * 6, 0xffffffffu, 0u);
    k_level_bounds<<<blocks_for(n), kThreads, 0, stream>>>(d_idx[cur].p, d_tbox.p, n, L, d_segb.p);
    const bool sort = L < sort_levels;
    k_level_keys<<<blocks_for(n), kThreads, 0, stream>>>(d_idx[cur].p, d_cen.p, d_segb.p, n, L, d_dead.p, sort ? d_keys[cur].p : nullptr);
    CKB(cudaGetLastError());
    if (sort) {
      cub::DoubleBuffer<unsigned long long> kb(d_keys[cur].p, d_keys[cur ^ 1].p);
      cub::DoubleBuffer<uint32_t> vb(d_idx[cur].p, d_idx[cur ^ 1].p);
      size_t tb = temp_bytes;
      CKB(cub::DeviceRadixSort::SortPairs(d_temp.p, tb, kb, vb, (int)n, 0, 32 + L, stream));  // stable, like the host's sort
      if (vb.Current() != d_idx[cur].p) cur ^= 1;  // (keys follow the same selector)
    }
  }
  k_finish_order<<<blocks_for(n), kThreads, 0, stream>>>(d_idx[cur].p, d_dead.p, n, d_order.p, d_dead_pos.p);
  CKB(cudaGetLastError());
  m.dead.resize(n);
  m.order.resize(n);
  std::vector<uint8_t> dead_pos(n);
  std::vector<uint32_t> idx_host(n);
  CKB(cudaMemcpyAsync(m.dead.data(), d_dead.p, n, cudaMemcpyDeviceToHost, stream));
  CKB(cudaMemcpyAsync(m.order.data(), d_order.p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
  CKB(cudaMemcpyAsync(dead_pos.data(), d_dead_pos.p, n, cudaMemcpyDeviceToHost, stream));
  CKB(cudaMemcpyAsync(idx_host.data(), d_idx[cur].p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
  CKB(cudaStreamSynchronize(stream));
  {
    int64_t nodes = 0, leaves = 0;
    int32_t depth = 0;
    std::unordered_map<uint64_t, std::pair<int64_t, int64_t>> memo;
    ref_counts(n, 0, nodes, leaves, depth, memo);
    m.ref_nodes = nodes, m.ref_leaves = leaves, m.ref_depth = depth;
  }
  const auto t2 = now();
  if (ref_only) {  // the caller builds the traversal tree itself (host SAH + collapse) from this dead mask and order
    m.ref_done = true;
    if (timing) {
      timing->upload_ms = ms_between(t0, t1), timing->ref_ms = ms_between(t1, t2), timing->total_ms = ms_between(t0, t2);
      timing->ref_levels = sort_levels + 1;
    }
    return;
  }

  // ---- step 2: structure on the host (live ranks only), boxes + nodes + triangle records on the device
  std::vector<uint32_t> R((size_t)n + 1);
  std::vector<uint32_t> live_tri;
  live_tri.reserve(n);
  R[0] = 0;
  for (uint32_t p = 0; p < n; p++) {
    R[p + 1] = R[p] + (dead_pos[p] ? 0u : 1u);
    if (!dead_pos[p]) live_tri.push_back(idx_host[p]);
  }
  m.live = (int64_t)R[n];
  Buf<float4> d_normals(n);
  k_normals<<<blocks_for(n), kThreads, 0, stream>>>(d_tris.p, d_order.p, n, d_normals.p);
  CKB(cudaGetLastError());
  if (m.live == 0) {
    // every triangle is unreachable in the reference: an empty root (no child bits) never reports a hit (as pt_build.cpp)
    m.nodes.assign(1, Node8());
    const uint32_t e = 127u | (127u << 8) | (127u << 16), one = 1u, inv = 0xffffffffu;
    float fe, fone, finv;
    memcpy(&fe, &e, 4), memcpy(&fone, &one, 4), memcpy(&finv, &inv, 4);
    m.nodes[0].q[0] = make_float4(0, 0, 0, fe);
    m.nodes[0].q[1] = make_float4(fone, 0.0f, 0.0f, 0.0f);
    m.nodes[0].q[2] = make_float4(finv, finv, finv, finv);
    m.nodes[0].q[3] = make_float4(finv, finv, 0.0f, 0.0f);
    m.nodes[0].q[4] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    Tri48 z;
    z.t[0] = z.t[1] = z.t[2] = make_float4(0, 0, 0, 0);
    m.tri48.assign(1, z);
    for (int a = 0; a < 3; a++) m.root_lo[a] = 1.0f, m.root_hi[a] = -1.0f;
    m.wide_depth = 0;
    m.normals.resize(n);
    CKB(cudaMemcpy(m.normals.data(), d_normals.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
    Buf<float4> d_nodes(5), d_tri48(3);
    CKB(cudaMemcpy(d_nodes.p, m.nodes.data(), 80, cudaMemcpyHostToDevice));
    CKB(cudaMemcpy(d_tri48.p, m.tri48.data(), 48, cudaMemcpyHostToDevice));
    out.nodes = d_nodes.take(), out.tris = d_tri48.take(), out.normals = d_normals.take(), out.device = device;
    m.built = true;
    return;
  }
  const Structure st = gen_structure(n, R);
  const uint32_t n_nodes = (uint32_t)st.nodes.size(), n_levels = (uint32_t)st.level_first.size() - 1;
  const auto t3 = now();
  uint32_t widest = 0;
  for (uint32_t l = 0; l < n_levels; l++) widest = std::max(widest, st.level_first[l + 1] - st.level_first[l]);
  size_t scan_bytes = 0;
  CKB(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)widest, stream));
  Arena a2(Arena::need(n_nodes, sizeof(WNode)) + Arena::need(live_tri.size(), 4) + Arena::need(st.quads.size(), 4) + Arena::need(n_nodes, sizeof(Box)) +
           4 * Arena::need(widest, 4) + Arena::need(scan_bytes, 1) + 8192);
  struct PW {
    WNode *p;
  } d_wn{a2.get<WNode>(n_nodes)};
  PU d_live{a2.get<uint32_t>(live_tri.size())};
  CKB(cudaMemcpyAsync(d_wn.p, st.nodes.data(), (size_t)n_nodes * sizeof(WNode), cudaMemcpyHostToDevice, stream));
  CKB(cudaMemcpyAsync(d_live.p, live_tri.data(), live_tri.size() * 4, cudaMemcpyHostToDevice, stream));
  if (!st.quads.empty()) {
    PU d_quads{a2.get<uint32_t>(st.quads.size())};
    CKB(cudaMemcpyAsync(d_quads.p, st.quads.data(), st.quads.size() * 4, cudaMemcpyHostToDevice, stream));
    k_pair_quads<<<blocks_for(st.quads.size()), kThreads, 0, stream>>>(d_quads.p, (uint32_t)st.quads.size(), d_live.p, d_tris.p);
    CKB(cudaGetLastError());
  }
  struct PX {
    Box *p;
  } d_box{a2.get<Box>(n_nodes)};
  Buf<float4> d_nodes((size_t)n_nodes * 5), d_tri48((size_t)std::max(st.n_tri_records, 1u) * 3);
  struct PF {
    float *p;
  } d_root{a2.get<float>(6)};
  struct PE {
    int *p;
  } d_err{a2.get<int>(1)};
  CKB(cudaMemsetAsync(d_err.p, 0, sizeof(int), stream));
  for (uint32_t l = n_levels; l-- > 0;) {
    const uint32_t first = st.level_first[l], count = st.level_first[l + 1] - first;
    k_boxes_level<<<blocks_for(count), kThreads, 0, stream>>>(d_wn.p, first, count, d_live.p, d_tris.p, d_box.p);
  }
  CKB(cudaGetLastError());
  PU d_sid[2] = {{a2.get<uint32_t>(widest)}, {a2.get<uint32_t>(widest)}}, d_cnt{a2.get<uint32_t>(widest)}, d_cbase{a2.get<uint32_t>(widest)};
  PB d_scan{a2.get<uint8_t>(scan_bytes)};
  CKB(cudaMemsetAsync(d_sid[0].p, 0, 4, stream));  // level 0: final position 0 holds structural node 0
  for (uint32_t l = 0; l < n_levels; l++) {
    const uint32_t first = st.level_first[l], count = st.level_first[l + 1] - first;
    k_inner_count<<<blocks_for(count), kThreads, 0, stream>>>(d_wn.p, d_sid[l & 1].p, count, d_cnt.p);
    size_t sb = scan_bytes;
    CKB(cub::DeviceScan::ExclusiveSum(d_scan.p, sb, d_cnt.p, d_cbase.p, (int)count, stream));
    k_emit_level<<<blocks_for(count), kThreads, 0, stream>>>(d_wn.p, d_sid[l & 1].p, first, count, d_cbase.p, st.level_first[l + 1], d_sid[(l + 1) & 1].p,
                                                        d_live.p, d_tris.p, d_order.p, d_box.p, d_nodes.p, d_tri48.p, d_root.p, d_err.p);
    CKB(cudaGetLastError());
  }
  int err = 0;
  float root[6];
  CKB(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CKB(cudaMemcpyAsync(root, d_root.p, sizeof(root), cudaMemcpyDeviceToHost, stream));
  CKB(cudaStreamSynchronize(stream));
  if (err) throw std::runtime_error("mesh extent too large for the quantised BVH frame");
  for (int a = 0; a < 3; a++) m.root_lo[a] = root[a], m.root_hi[a] = root[3 + a];
  m.wide_depth = (int32_t)n_levels - 1;
  const auto t4 = now();

  // ---- host copies (ptc_scene_mesh_info, replicas of ptc_multi_create)
  m.nodes.resize(n_nodes);
  m.tri48.resize(std::max(st.n_tri_records, 1u));
  m.normals.resize(n);
  CKB(cudaMemcpyAsync(m.nodes.data(), d_nodes.p, (size_t)n_nodes * 80, cudaMemcpyDeviceToHost, stream));
  CKB(cudaMemcpyAsync(m.tri48.data(), d_tri48.p, m.tri48.size() * 48, cudaMemcpyDeviceToHost, stream));
  CKB(cudaMemcpyAsync(m.normals.data(), d_normals.p, (size_t)n * 16, cudaMemcpyDeviceToHost, stream));
  CKB(cudaStreamSynchronize(stream));
  out.nodes = d_nodes.take(), out.tris = d_tri48.take(), out.normals = d_normals.take(), out.device = device;
  m.built = true;
  const auto t5 = now();
  if (timing) {
    timing->upload_ms = ms_between(t0, t1), timing->ref_ms = ms_between(t1, t2), timing->structure_ms = ms_between(t2, t3);
    timing->emit_ms = ms_between(t3, t4), timing->download_ms = ms_between(t4, t5), timing->total_ms = ms_between(t0, t5);
    timing->ref_levels = sort_levels + 1;
  }
}

}  // namespace pt
