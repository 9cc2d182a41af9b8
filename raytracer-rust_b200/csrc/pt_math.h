// pt_math.h — L0 math of the hot path with the reference's exact semantics.
//   Vec3   src/vec3.rs:6-129      Ray  src/ray.rs:4-17      Color ops src/color.rs:32-85
//   glam::Mat4 * Vec4 as used by src/objects/cube.rs:63-76,126-134 and src/mesh/mesh_object.rs:264-310
#pragma once
#include "pt_hd.h"

namespace pt {

constexpr float kEps = 1e-4f;  // EPSILON: renderer.rs:17, material.rs:8, tungsten/materials.rs:9
constexpr float kPi = 3.14159265358979323846f;

struct V3 {
  float x, y, z;
};

PT_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
PT_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PT_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PT_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
PT_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }  // Color * Color
PT_HD V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }  // Color / Color
PT_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
PT_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
// vec3.rs:17-19: (x*x' + y*y') + z*z', three roundings for the products, two for the sums
PT_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD V3 cross(V3 a, V3 b) {  // vec3.rs:21-27
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
PT_HD float length_squared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
PT_HD float length(V3 a) { return sqrtf(length_squared(a)); }
PT_HD V3 normalized(V3 a) {  // vec3.rs:37-44: returned unchanged when shorter than EPSILON
  float len = length(a);
  if (len < kEps) return a;
  return a * (1.0f / len);
}
PT_HD bool near_zero(V3 a) {  // vec3.rs:63-66
  const float S = 1e-8f;
  return fabsf(a.x) < S && fabsf(a.y) < S && fabsf(a.z) < S;
}
PT_HD bool has_nan(V3 a) { return isnan_f(a.x) || isnan_f(a.y) || isnan_f(a.z); }
PT_HD bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
PT_HD V3 reflect_plain(V3 v, V3 n) { return v - n * 2.0f * dot(v, n); }  // vec3.rs:68-70
PT_HD V3 to_world(V3 local, V3 n) {                                       // vec3.rs:72-81
  V3 up = fabsf(n.z) < 0.999f ? v3(0, 0, 1) : v3(0, 1, 0);
  V3 tangent = normalized(cross(n, up));
  V3 bitangent = cross(n, tangent);
  return tangent * local.x + bitangent * local.y + n * local.z;
}
PT_HD V3 sqrt3(V3 a) { return V3{sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)}; }
PT_HD V3 splat(float v) { return V3{v, v, v}; }
PT_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// f32::signum: +1 for +0 and positives, -1 for -0 and negatives, NaN stays NaN
PT_HD float signum(float v) {
  if (isnan_f(v)) return v;
  return copysignf(1.0f, v);
}

struct Ray {
  V3 o, d;
};
PT_HD V3 ray_at(const Ray &r, float t) { return r.o + r.d * t; }  // ray.rs:9-11
PT_HD Ray ray_new(V3 o, V3 d) { return Ray{o, normalized(d)}; }    // ray.rs:12-17

// Column-major 4x4 (glam::Mat4::to_cols_array).  m[c*4 + r].
// glam Mat4 * Vec4, scalar and sse2 paths alike: ((x_axis*v.x + y_axis*v.y) + z_axis*v.z) + w_axis*v.w
PT_HD V3 mat_point(const float *m, V3 p) {  // w = 1; only xyz of the result are used by the reference
  V3 r;
  r.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * 1.0f;
  r.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * 1.0f;
  r.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14] * 1.0f;
  return r;
}
PT_HD V3 mat_vector(const float *m, V3 v) {  // w = 0
  V3 r;
  r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * 0.0f;
  r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * 0.0f;
  r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * 0.0f;
  return r;
}
// transpose(m) * (v, 0): rows of the product are the columns of m
PT_HD V3 mat_t_vector(const float *m, V3 v) {
  V3 r;
  r.x = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * 0.0f;
  r.y = ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7] * 0.0f;
  r.z = ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11] * 0.0f;
  return r;
}


// The same three products with the matrix held as four 16-byte columns (one LDG.128 each instead of sixteen scalar
// loads); identical operation order, identical bits.
struct M4 {
  float4 c0, c1, c2, c3;
};
PT_HD M4 load_m4(const float *m) {  // m must be 16-byte aligned (DObject::f is)
  const float4 *p = reinterpret_cast<const float4 *>(m);
  return M4{ld4(p), ld4(p + 1), ld4(p + 2), ld4(p + 3)};
}
PT_HD V3 mat_point(const M4 &m, V3 p) {
  V3 r;
  r.x = ((m.c0.x * p.x + m.c1.x * p.y) + m.c2.x * p.z) + m.c3.x * 1.0f;
  r.y = ((m.c0.y * p.x + m.c1.y * p.y) + m.c2.y * p.z) + m.c3.y * 1.0f;
  r.z = ((m.c0.z * p.x + m.c1.z * p.y) + m.c2.z * p.z) + m.c3.z * 1.0f;
  return r;
}
PT_HD V3 mat_vector(const M4 &m, V3 v) {
  V3 r;
  r.x = ((m.c0.x * v.x + m.c1.x * v.y) + m.c2.x * v.z) + m.c3.x * 0.0f;
  r.y = ((m.c0.y * v.x + m.c1.y * v.y) + m.c2.y * v.z) + m.c3.y * 0.0f;
  r.z = ((m.c0.z * v.x + m.c1.z * v.y) + m.c2.z * v.z) + m.c3.z * 0.0f;
  return r;
}
// ---- exact packed FP32 (sm_100a fma.rn.f32x2, SASS FFMA2: one issue slot for two IEEE operations) -------------------
// The extend stage is bound by instruction issue, not by the FMA pipe (tools/micro/ffma2_bench.cu), so the unfused
// multiplies and adds of the reference's arithmetic are issued two at a time:
//     mul2: fma(a, b, -0.0) == RN(a * b)   (adding -0 changes no product, signed zeros included)
//     add2: fma(a, 1.0, c)  == RN(a + c)   (a * 1 is exact)
//     sub2: fma(b, -1.0, a) == RN(a - b)
// each result rounded once, exactly like FMUL / FADD.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even
// under --fmad=false, so the operations are spelled as FMAs whose constants it cannot see (constant memory), which it
// has no licence to merge.  The host build (hostsim) runs the scalar statements the packed ones replace.
#if defined(__CUDACC__)
__constant__ float kPackConst[4] = {-0.0f, 1.0f, -1.0f, 0.0f};
struct F2 {
  float x, y;
};
__device__ __forceinline__ F2 fma2(float ax, float ay, float bx, float by, float cx, float cy) {
  F2 r;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(ax), "f"(ay), "f"(bx), "f"(by), "f"(cx), "f"(cy));
  return r;
}
__device__ __forceinline__ F2 mul2(float ax, float ay, float bx, float by) { return fma2(ax, ay, bx, by, kPackConst[0], kPackConst[0]); }
__device__ __forceinline__ F2 add2(float ax, float ay, float cx, float cy) { return fma2(ax, ay, kPackConst[1], kPackConst[1], cx, cy); }
__device__ __forceinline__ F2 sub2(float ax, float ay, float bx, float by) { return fma2(bx, by, kPackConst[2], kPackConst[2], ax, ay); }
#endif

// mat_point(m, o) and mat_vector(m, d) of the same matrix in one go (what Cube::hit and Mesh::hit start with): x and y
// of each product as packed pairs, the two z components as a third pair.  Same operations in the same order as the two
// functions above, component by component.
template <bool PACKED = true>
PT_HD void mat_point_vector(const M4 &m, V3 o, V3 d, V3 &po, V3 &pd) {
#if defined(__CUDA_ARCH__)
  if (!PACKED) {  // k_extend_post waits on memory, not on issue slots: there the packing only adds moves
    po = mat_point(m, o);
    pd = mat_vector(m, d);
    return;
  }
  F2 t, u;
  t = mul2(m.c0.x, m.c0.y, o.x, o.x), u = mul2(m.c1.x, m.c1.y, o.y, o.y), t = add2(t.x, t.y, u.x, u.y);
  u = mul2(m.c2.x, m.c2.y, o.z, o.z), t = add2(t.x, t.y, u.x, u.y);
  t = add2(t.x, t.y, m.c3.x, m.c3.y);  // + c3 * 1.0f
  po.x = t.x, po.y = t.y;
  t = mul2(m.c0.x, m.c0.y, d.x, d.x), u = mul2(m.c1.x, m.c1.y, d.y, d.y), t = add2(t.x, t.y, u.x, u.y);
  u = mul2(m.c2.x, m.c2.y, d.z, d.z), t = add2(t.x, t.y, u.x, u.y);
  u = mul2(m.c3.x, m.c3.y, kPackConst[3], kPackConst[3]), t = add2(t.x, t.y, u.x, u.y);  // + c3 * 0.0f
  pd.x = t.x, pd.y = t.y;
  // (po.z, pd.z)
  t = mul2(o.x, d.x, m.c0.z, m.c0.z), u = mul2(o.y, d.y, m.c1.z, m.c1.z), t = add2(t.x, t.y, u.x, u.y);
  u = mul2(o.z, d.z, m.c2.z, m.c2.z), t = add2(t.x, t.y, u.x, u.y);
  po.z = t.x + m.c3.z * 1.0f;
  pd.z = t.y + m.c3.z * 0.0f;
#else
  po = mat_point(m, o);
  pd = mat_vector(m, d);
#endif
}
// mat_point alone, x and y packed
PT_HD V3 mat_point_packed(const M4 &m, V3 p) {
#if defined(__CUDA_ARCH__)
  F2 t, u;
  t = mul2(m.c0.x, m.c0.y, p.x, p.x), u = mul2(m.c1.x, m.c1.y, p.y, p.y), t = add2(t.x, t.y, u.x, u.y);
  u = mul2(m.c2.x, m.c2.y, p.z, p.z), t = add2(t.x, t.y, u.x, u.y);
  t = add2(t.x, t.y, m.c3.x, m.c3.y);
  V3 r;
  r.x = t.x, r.y = t.y;
  r.z = ((m.c0.z * p.x + m.c1.z * p.y) + m.c2.z * p.z) + m.c3.z * 1.0f;
  return r;
#else
  return mat_point(m, p);
#endif
}

PT_HD V3 mat_t_vector(const M4 &m, V3 v) {
  V3 r;
  r.x = ((m.c0.x * v.x + m.c0.y * v.y) + m.c0.z * v.z) + m.c0.w * 0.0f;
  r.y = ((m.c1.x * v.x + m.c1.y * v.y) + m.c1.z * v.z) + m.c1.w * 0.0f;
  r.z = ((m.c2.x * v.x + m.c2.y * v.y) + m.c2.z * v.z) + m.c2.w * 0.0f;
  return r;
}

}  // namespace pt
