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
PT_HD V3 mat_t_vector(const M4 &m, V3 v) {
  V3 r;
  r.x = ((m.c0.x * v.x + m.c0.y * v.y) + m.c0.z * v.z) + m.c0.w * 0.0f;
  r.y = ((m.c1.x * v.x + m.c1.y * v.y) + m.c1.z * v.z) + m.c1.w * 0.0f;
  r.z = ((m.c2.x * v.x + m.c2.y * v.y) + m.c2.z * v.z) + m.c2.w * 0.0f;
  return r;
}

}  // namespace pt
