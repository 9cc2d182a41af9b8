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
// Box tests here are conservative (boxes padded at build time), so they may fuse and use an approximate reciprocal.
//
// The traversal is written as single steps (one node, one triangle) over an explicit state, so that the same code
// serves the straight-line loop below (parity hooks, hostsim) and the persistent-warp kernel, which interleaves the
// steps of 32 rays and refills finished lanes (ptcore.cu: k_traverse).
#pragma once
#include "pt_math.h"
#include "pt_types.h"

namespace pt {

struct MeshHit {
  float t;
  uint32_t tri;    // original triangle index, 0xffffffff = none
  uint32_t order;  // reference DFS position of the winner
};

// reciprocal for slab tests only (never for anything the parity bar looks at)
PT_HD float box_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}

// Byte j of `packed` as the float 1 + byte * 2^-15, built by ONE byte permute (no int->float conversion: I2F runs at
// an eighth of the FP32 rate on sm_100 and there are 48 of them per node).  The slab arithmetic absorbs the affine
// map: q * adj + org == m * (adj * 2^15) + (org - adj * 2^15) with m = 1 + q * 2^-15; the extra rounding error is
// ~2^-17 of the node's t-extent, far inside the 2^-9 padding the builder gives every frame.
template <int J>
PT_HD float byte_as_unit_float(uint32_t packed) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(__byte_perm(packed, 0x3F800000u, 0x7604u | (uint32_t)(J << 4)));
#else
  return u2f(0x3F800000u | (((packed >> (8 * J)) & 0xffu) << 8));
#endif
}

// (r0, r1) = (a0, a1) * b + c with b, c scalars: ONE packed FFMA2 on the device (conservative box arithmetic only; the
// host build computes the same two fused products)
PT_HD void fma2_bcast(float &r0, float &r1, float a0, float a1, float b, float c) {
#if defined(__CUDA_ARCH__)
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%4}; mov.b64 rc, {%5,%5}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(r0), "=f"(r1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
#else
  r0 = fmaf(a0, b, c);
  r1 = fmaf(a1, b, c);
#endif
}

struct TravState {
  V3 o, d;             // object-space ray
  float idx, idy, idz;  // 1/d for the slab tests
  float t_min, best_t;
  uint32_t best_order, best_tri;
  uint32_t octinv;
  uint2 ng;  // node group: x = first child index, y = hit bits [24,32) | imask
  uint2 tg;  // triangle group: x = first triangle, y = hit bits [0,24)
  uint2 tg2; // second triangle group, filled by a node step taken while `tg` still has triangles pending (tg2 != 0 => tg != 0)
};

PT_HD void trav_begin(TravState &s, V3 o, V3 d, float t_min, float t_max) {
  s.o = o;
  s.d = d;
  s.idx = box_rcp(d.x), s.idy = box_rcp(d.y), s.idz = box_rcp(d.z);
  s.t_min = t_min;
  s.best_t = t_max;
  s.best_order = 0u;
  s.best_tri = 0xffffffffu;
  const uint32_t sx = f2u(d.x) >> 31, sy = f2u(d.y) >> 31, sz = f2u(d.z) >> 31;
  s.octinv = 7u - (sx | (sy << 1) | (sz << 2));
  s.ng = make_uint2(0u, 0x80000000u);
  s.tg = make_uint2(0u, 0u);
  s.tg2 = make_uint2(0u, 0u);
}
PT_HD bool trav_has_node(const TravState &s) { return s.ng.y > 0x00FFFFFFu; }
PT_HD bool trav_has_tri(const TravState &s) { return s.tg.y != 0u; }

// Pops the nearest pending child of the node group, pushes the rest, tests the child's 8 boxes.
// Precondition: trav_has_node(s) && s.tg2.y == 0 (the new leaf hits go to `tg`, or to `tg2` while `tg` is still busy, so
// a lane can walk on to its next node while it works through the triangles of the previous one).
template <bool COUNT>
PT_HD void trav_node(const DMesh &m, TravState &s, uint2 *stack, int &sp, TraversalCounters *ctr) {
  const uint32_t hits = s.ng.y, imask = hits & 0xffu;
  const int bit = 31 - clz32(hits);
  s.ng.y &= ~(1u << bit);
  if (s.ng.y > 0x00FFFFFFu) stack[sp++] = s.ng;
  const uint32_t slot = (uint32_t)(bit - 24) ^ s.octinv;
  const uint32_t rel = (uint32_t)popc32(imask & ~(0xFFFFFFFFu << slot));
  const float4 *np = m.nodes + (size_t)(s.ng.x + rel) * 5;
  const float4 q0 = ldg4(np), q1 = ldg4(np + 1), q2 = ldg4(np + 2), q3 = ldg4(np + 3), q4 = ldg4(np + 4);
  if (COUNT) ctr->nodes++;

  const uint32_t e = f2u(q0.w);
  // scale * 2^15 folded into the exponent byte (the builder keeps biased exponents <= 254 - 15)
  const float adjx = u2f(((e & 0xffu) + 15u) << 23) * s.idx;
  const float adjy = u2f((((e >> 8) & 0xffu) + 15u) << 23) * s.idy;
  const float adjz = u2f((((e >> 16) & 0xffu) + 15u) << 23) * s.idz;
  const float orgx = (q0.x - s.o.x) * s.idx - adjx, orgy = (q0.y - s.o.y) * s.idy - adjy, orgz = (q0.z - s.o.z) * s.idz - adjz;

  // near / far planes by ray octant
  const bool sx = (f2u(s.d.x) >> 31) != 0u, sy = (f2u(s.d.y) >> 31) != 0u, sz = (f2u(s.d.z) >> 31) != 0u;
  const uint32_t lox0 = f2u(q2.x), lox1 = f2u(q2.y), loy0 = f2u(q2.z), loy1 = f2u(q2.w);
  const uint32_t loz0 = f2u(q3.x), loz1 = f2u(q3.y), hix0 = f2u(q3.z), hix1 = f2u(q3.w);
  const uint32_t hiy0 = f2u(q4.x), hiy1 = f2u(q4.y), hiz0 = f2u(q4.z), hiz1 = f2u(q4.w);
  const uint32_t nx[2] = {sx ? hix0 : lox0, sx ? hix1 : lox1}, fx[2] = {sx ? lox0 : hix0, sx ? lox1 : hix1};
  const uint32_t ny[2] = {sy ? hiy0 : loy0, sy ? hiy1 : loy1}, fy[2] = {sy ? loy0 : hiy0, sy ? loy1 : hiy1};
  const uint32_t nz[2] = {sz ? hiz0 : loz0, sz ? hiz1 : loz1}, fz[2] = {sz ? loz0 : hiz0, sz ? loz1 : hiz1};
  const uint32_t meta[2] = {f2u(q1.z), f2u(q1.w)};
  const uint32_t octinv4 = s.octinv * 0x01010101u;

  uint32_t hitmask = 0u;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t meta4 = meta[h];
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xffu;
    const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    // Two children per instruction: sm_100a's packed fma.rn.f32x2 (SASS FFMA2) takes one issue slot for two FMAs, and
    // this stage is bound by issue slots, not by the FMA pipe (tools/micro/ffma2_bench.cu: FP32 work mixed with integer
    // work runs 24 % faster packed).  The six slab distances of children J and J+1 are six FFMA2 instead of twelve FFMA.
#define PT_CHILD2(J)                                                                                                  \
  {                                                                                                                   \
    float tnx0, tnx1, tny0, tny1, tnz0, tnz1, tfx0, tfx1, tfy0, tfy1, tfz0, tfz1;                                     \
    fma2_bcast(tnx0, tnx1, byte_as_unit_float<J>(nx[h]), byte_as_unit_float<J + 1>(nx[h]), adjx, orgx);               \
    fma2_bcast(tny0, tny1, byte_as_unit_float<J>(ny[h]), byte_as_unit_float<J + 1>(ny[h]), adjy, orgy);               \
    fma2_bcast(tnz0, tnz1, byte_as_unit_float<J>(nz[h]), byte_as_unit_float<J + 1>(nz[h]), adjz, orgz);               \
    fma2_bcast(tfx0, tfx1, byte_as_unit_float<J>(fx[h]), byte_as_unit_float<J + 1>(fx[h]), adjx, orgx);               \
    fma2_bcast(tfy0, tfy1, byte_as_unit_float<J>(fy[h]), byte_as_unit_float<J + 1>(fy[h]), adjy, orgy);               \
    fma2_bcast(tfz0, tfz1, byte_as_unit_float<J>(fz[h]), byte_as_unit_float<J + 1>(fz[h]), adjz, orgz);               \
    /* fmaxf / fminf drop NaN operands (0 * inf from axis-parallel rays): the slab then does not constrain */         \
    const float tn0 = fmaxf(fmaxf(tnx0, tny0), fmaxf(tnz0, s.t_min)), tf0 = fminf(fminf(tfx0, tfy0), fminf(tfz0, s.best_t)); \
    const float tn1 = fmaxf(fmaxf(tnx1, tny1), fmaxf(tnz1, s.t_min)), tf1 = fminf(fminf(tfx1, tfy1), fminf(tfz1, s.best_t)); \
    if (tn0 <= tf0) hitmask |= ((child_bits4 >> (8 * (J))) & 0xffu) << ((bit_index4 >> (8 * (J))) & 0xffu);           \
    if (tn1 <= tf1) hitmask |= ((child_bits4 >> (8 * (J + 1))) & 0xffu) << ((bit_index4 >> (8 * (J + 1))) & 0xffu);   \
  }
    PT_CHILD2(0) PT_CHILD2(2)
#undef PT_CHILD
  }
  s.ng.x = f2u(q1.x);
  s.ng.y = (hitmask & 0xFF000000u) | (e >> 24);
  const uint2 leaf_hits = make_uint2(f2u(q1.y), hitmask & 0x00FFFFFFu);
  if (s.tg.y == 0u) s.tg = leaf_hits;
  else s.tg2 = leaf_hits;
}

// Tests ONE pending triangle of the triangle group.  Precondition: trav_has_tri(s).
template <bool COUNT>
PT_HD void trav_tri(const DMesh &m, TravState &s, TraversalCounters *ctr) {
  const int bit = 31 - clz32(s.tg.y);
  const float4 *tp = m.tris + (size_t)(s.tg.x + (uint32_t)bit) * 3;
  s.tg.y &= ~(1u << bit);
  if (s.tg.y == 0u) {
    s.tg = s.tg2;
    s.tg2 = make_uint2(0u, 0u);
  }
  const float4 t0 = ldg4(tp), t1 = ldg4(tp + 1), t2 = ldg4(tp + 2);
  if (COUNT) ctr->tris++;
  // bvh.rs:94-116, same operation order, no fusing
  const V3 v0 = v3(t0.x, t0.y, t0.z), edge1 = v3(t1.x, t1.y, t1.z), edge2 = v3(t2.x, t2.y, t2.z);
  const V3 h = cross(s.d, edge2);
  const float a = dot(edge1, h);
  if (fabsf(a) < kEps) return;
  const float f = 1.0f / a;
  const V3 sv = s.o - v0;
  const float u = f * dot(sv, h);
  if (!(0.0f <= u && u <= 1.0f)) return;
  const V3 q = cross(sv, edge1);
  const float v = f * dot(s.d, q);
  if (v < 0.0f || u + v > 1.0f) return;
  const float t = f * dot(edge2, q);
  if (!(t > s.t_min)) return;
  const uint32_t order = f2u(t1.w);
  if (t < s.best_t || (t == s.best_t && order < s.best_order)) {
    s.best_t = t;
    s.best_order = order;
    s.best_tri = f2u(t0.w);
  }
}

// Straight-line closest hit: the steps above until nothing is pending.
template <bool COUNT>
PT_HD bool bvh8_closest(const DMesh &m, V3 o, V3 d, float t_min, float t_max, MeshHit &out, TraversalCounters *ctr) {
  TravState s;
  trav_begin(s, o, d, t_min, t_max);
  uint2 stack[kTraversalStack];
  int sp = 0;
  for (;;) {
    if (trav_has_tri(s)) {
      trav_tri<COUNT>(m, s, ctr);
    } else if (trav_has_node(s)) {
      trav_node<COUNT>(m, s, stack, sp, ctr);
    } else {
      if (sp == 0) break;
      s.ng = stack[--sp];
    }
  }
  if (s.best_tri == 0xffffffffu) return false;
  out.t = s.best_t;
  out.tri = s.best_tri;
  out.order = s.best_order;
  return true;
}

}  // namespace pt
