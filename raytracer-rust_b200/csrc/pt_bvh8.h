// pt_bvh8.h — software traversal of the compressed 8-wide BVH + the reference's triangle test.
//
// Replaces BVHNode::intersect_recursive (src/acceleration/bvh.rs:78-170) and Aabb::intersect
// (src/acceleration/aabb.rs:27-45).  What must stay identical to the reference is the RESULT of the closest-hit
// query, not the tree:
//   * a triangle is accepted by exactly the arithmetic of bvh.rs:94-118 (Moeller-Trumbore, |a| < 1e-4 cull,
//     u in [0,1], v >= 0, u+v <= 1, t in the OPEN interval (t_min, t_max)), unfused;
//   * on equal t the triangle that comes first in the reference's depth-first leaf order wins (left subtree
//     before right, bvh.rs:142-156; leaf loop in index order with a strict `t < t_max`, bvh.rs:118,135).  That
//     order is precomputed on the host and stored in the triangle record, so the wide tree may visit in any order;
//   * triangles under a zero-extent reference node can never be hit (aabb.rs:40 `t_max <= t_min`): the host
//     builder leaves them out of this tree.
// Box tests here are conservative (boxes padded at build time), so they may fuse.
#pragma once
#include "pt_math.h"
#include "pt_types.h"

namespace pt {

struct MeshHit {
  float t;
  uint32_t tri;    // original triangle index
  uint32_t order;  // reference DFS position of the winner
};

template <bool COUNT>
PT_HD bool bvh8_closest(const DMesh &m, V3 o, V3 d, float t_min, float t_max, MeshHit &out, TraversalCounters *ctr) {
  float best_t = t_max;
  uint32_t best_order = 0u, best_tri = 0xffffffffu;

  const float idx = 1.0f / d.x, idy = 1.0f / d.y, idz = 1.0f / d.z;
  const uint32_t sx = f2u(d.x) >> 31, sy = f2u(d.y) >> 31, sz = f2u(d.z) >> 31;
  const uint32_t octinv = 7u - (sx | (sy << 1) | (sz << 2));
  const uint32_t octinv4 = octinv * 0x01010101u;

  uint2 stack[kTraversalStack];
  int sp = 0;
  uint2 ng = make_uint2(0u, 0x80000000u);  // node group: x = first child index, y = hit bits [24,32) | imask
  uint2 tg = make_uint2(0u, 0u);           // triangle group: x = first triangle, y = hit bits [0,24)

  for (;;) {
    if (ng.y > 0x00FFFFFFu) {
      const uint32_t hits = ng.y, imask = hits & 0xffu;
      const int bit = 31 - clz32(hits);
      ng.y &= ~(1u << bit);
      if (ng.y > 0x00FFFFFFu) stack[sp++] = ng;
      const uint32_t slot = (uint32_t)(bit - 24) ^ octinv;
      const uint32_t rel = (uint32_t)popc32(imask & ~(0xFFFFFFFFu << slot));
      const float4 *np = m.nodes + (size_t)(ng.x + rel) * 5;
      const float4 q0 = ldg4(np), q1 = ldg4(np + 1), q2 = ldg4(np + 2), q3 = ldg4(np + 3), q4 = ldg4(np + 4);
      if (COUNT) ctr->nodes++;

      const uint32_t e = f2u(q0.w);
      const float adjx = u2f((e & 0xffu) << 23) * idx;
      const float adjy = u2f(((e >> 8) & 0xffu) << 23) * idy;
      const float adjz = u2f(((e >> 16) & 0xffu) << 23) * idz;
      const float orgx = (q0.x - o.x) * idx, orgy = (q0.y - o.y) * idy, orgz = (q0.z - o.z) * idz;

      // near / far planes by ray octant
      const uint32_t lox0 = f2u(q2.x), lox1 = f2u(q2.y), loy0 = f2u(q2.z), loy1 = f2u(q2.w);
      const uint32_t loz0 = f2u(q3.x), loz1 = f2u(q3.y), hix0 = f2u(q3.z), hix1 = f2u(q3.w);
      const uint32_t hiy0 = f2u(q4.x), hiy1 = f2u(q4.y), hiz0 = f2u(q4.z), hiz1 = f2u(q4.w);
      const uint32_t nx[2] = {sx ? hix0 : lox0, sx ? hix1 : lox1}, fx[2] = {sx ? lox0 : hix0, sx ? lox1 : hix1};
      const uint32_t ny[2] = {sy ? hiy0 : loy0, sy ? hiy1 : loy1}, fy[2] = {sy ? loy0 : hiy0, sy ? loy1 : hiy1};
      const uint32_t nz[2] = {sz ? hiz0 : loz0, sz ? hiz1 : loz1}, fz[2] = {sz ? loz0 : hiz0, sz ? loz1 : hiz1};
      const uint32_t meta[2] = {f2u(q1.z), f2u(q1.w)};

      uint32_t hitmask = 0u;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const uint32_t meta4 = meta[h];
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xffu;
        const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int sh = 8 * j;
          const float tnx = fmaf((float)((nx[h] >> sh) & 0xffu), adjx, orgx);
          const float tny = fmaf((float)((ny[h] >> sh) & 0xffu), adjy, orgy);
          const float tnz = fmaf((float)((nz[h] >> sh) & 0xffu), adjz, orgz);
          const float tfx = fmaf((float)((fx[h] >> sh) & 0xffu), adjx, orgx);
          const float tfy = fmaf((float)((fy[h] >> sh) & 0xffu), adjy, orgy);
          const float tfz = fmaf((float)((fz[h] >> sh) & 0xffu), adjz, orgz);
          // fmaxf / fminf drop NaN operands (0 * inf from axis-parallel rays): the slab then does not constrain
          const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, t_min));
          const float tf = fminf(fminf(tfx, tfy), fminf(tfz, best_t));
          if (tn <= tf) hitmask |= ((child_bits4 >> sh) & 0xffu) << ((bit_index4 >> sh) & 0xffu);
        }
      }
      ng.x = f2u(q1.x);
      ng.y = (hitmask & 0xFF000000u) | (e >> 24);
      tg.x = f2u(q1.y);
      tg.y = hitmask & 0x00FFFFFFu;
    } else {
      tg = ng;
      ng = make_uint2(0u, 0u);
    }

    while (tg.y != 0u) {
      const int bit = 31 - clz32(tg.y);
      tg.y &= ~(1u << bit);
      const float4 *tp = m.tris + (size_t)(tg.x + (uint32_t)bit) * 3;
      const float4 t0 = ldg4(tp), t1 = ldg4(tp + 1), t2 = ldg4(tp + 2);
      if (COUNT) ctr->tris++;
      // bvh.rs:94-116, same operation order, no fusing
      const V3 v0 = v3(t0.x, t0.y, t0.z), edge1 = v3(t1.x, t1.y, t1.z), edge2 = v3(t2.x, t2.y, t2.z);
      const V3 h = cross(d, edge2);
      const float a = dot(edge1, h);
      if (fabsf(a) < kEps) continue;
      const float f = 1.0f / a;
      const V3 s = o - v0;
      const float u = f * dot(s, h);
      if (!(0.0f <= u && u <= 1.0f)) continue;
      const V3 q = cross(s, edge1);
      const float v = f * dot(d, q);
      if (v < 0.0f || u + v > 1.0f) continue;
      const float t = f * dot(edge2, q);
      if (!(t > t_min)) continue;
      const uint32_t order = f2u(t1.w);
      if (t < best_t || (t == best_t && order < best_order)) {
        best_t = t;
        best_order = order;
        best_tri = f2u(t0.w);
      }
    }

    if (ng.y <= 0x00FFFFFFu) {
      if (sp == 0) break;
      ng = stack[--sp];
    }
  }
  if (best_tri == 0xffffffffu) return false;
  out.t = best_t;
  out.tri = best_tri;
  out.order = best_order;
  return true;
}

}  // namespace pt
