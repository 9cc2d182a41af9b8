// oracle.cpp — CPU restatement of the raytracer-rust hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Build with -ffp-contract=off and without -ffast-math: rustc never contracts a*b+c into an FMA and the
// reference's Cargo.toml has no target-cpu flags, so every fp32 operation below rounds exactly once.
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

constexpr float EPSILON = 1e-4f;  // renderer.rs:17, material.rs:8, tungsten/materials.rs:9
constexpr float PI = 3.14159265358979323846f;

// Rust's f32::min / f32::max ignore a NaN operand, like fminf / fmaxf.
inline float rmin(float a, float b) { return std::fmin(a, b); }
inline float rmax(float a, float b) { return std::fmax(a, b); }

// ---------------------------------------------------------------- vec3.rs
struct Vec3 {
  float x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  float dot(Vec3 o) const { return x * o.x + y * o.y + z * o.z; }  // vec3.rs:17-19
  Vec3 cross(Vec3 o) const {                                         // vec3.rs:21-27
    return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
  }
  float length_squared() const { return x * x + y * y + z * z; }  // vec3.rs:29-31
  float length() const { return std::sqrt(length_squared()); }    // vec3.rs:33-35
  Vec3 operator+(Vec3 o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
  Vec3 operator-(Vec3 o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
  Vec3 operator*(float s) const { return Vec3(x * s, y * s, z * s); }
  Vec3 operator-() const { return Vec3(-x, -y, -z); }
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  Vec3 normalized() const {  // vec3.rs:37-44: unchanged if shorter than EPSILON
    float len = length();
    if (len < EPSILON) return *this;
    return *this * (1.0f / len);
  }
  bool near_zero() const {  // vec3.rs:63-66
    const float S = 1e-8f;
    return std::fabs(x) < S && std::fabs(y) < S && std::fabs(z) < S;
  }
  Vec3 reflect(Vec3 n) const { return *this - n * 2.0f * dot(n); }  // vec3.rs:68-70
  bool has_nan() const { return std::isnan(x) || std::isnan(y) || std::isnan(z); }
  bool is_zero() const { return x == 0.0f && y == 0.0f && z == 0.0f; }
};

// vec3.rs:117-129 — Vec3 / f32 panics if |s| < EPSILON.  The oracle cannot unwind into C; it
// reports the condition through a flag the tests can read and carries on with IEEE division.
std::atomic<int> g_div_panics{0};
inline Vec3 vdiv(Vec3 v, float s) {
  if (std::fabs(s) < EPSILON) g_div_panics.fetch_add(1, std::memory_order_relaxed);
  return Vec3(v.x / s, v.y / s, v.z / s);
}

Vec3 to_world(Vec3 local, Vec3 normal) {  // vec3.rs:72-81
  Vec3 up = std::fabs(normal.z) < 0.999f ? Vec3(0, 0, 1) : Vec3(0, 1, 0);
  Vec3 tangent = normal.cross(up).normalized();
  Vec3 bitangent = normal.cross(tangent);
  return tangent * local.x + bitangent * local.y + normal * local.z;
}

// ---------------------------------------------------------------- color.rs
struct Color {
  float r, g, b;
  Color() : r(0), g(0), b(0) {}
  Color(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
  static Color splat(float v) { return Color(v, v, v); }
  Color operator+(Color o) const { return Color(r + o.r, g + o.g, b + o.b); }
  Color operator-(Color o) const { return Color(r - o.r, g - o.g, b - o.b); }
  Color operator*(Color o) const { return Color(r * o.r, g * o.g, b * o.b); }
  Color operator*(float s) const { return Color(r * s, g * s, b * s); }
  Color operator/(Color o) const { return Color(r / o.r, g / o.g, b / o.b); }
  Color operator/(float s) const { return Color(r / s, g / s, b / s); }
  Color sqrt() const { return Color(std::sqrt(r), std::sqrt(g), std::sqrt(b)); }
};

// ---------------------------------------------------------------- ray.rs
struct Ray {
  Vec3 origin, direction;
  Ray() {}
  Ray(Vec3 o, Vec3 d) : origin(o), direction(d.normalized()) {}  // ray.rs:12-17
  static Ray raw(Vec3 o, Vec3 d) {                               // for caller-provided rays
    Ray r;
    r.origin = o;
    r.direction = d;
    return r;
  }
  Vec3 at(float t) const { return origin + direction * t; }  // ray.rs:9-11
};

// ---------------------------------------------------------------- glam Mat4 (column-major), used by cube.rs / mesh_object.rs
struct Mat4 {
  float c[4][4];  // c[col][row]
  // glam Mat4 * Vec4 (sse2 and scalar paths agree): ((x_axis*v.x + y_axis*v.y) + z_axis*v.z) + w_axis*v.w
  void mul(float vx, float vy, float vz, float vw, float out[4]) const {
    for (int i = 0; i < 4; i++) {
      float r = c[0][i] * vx;
      r = r + c[1][i] * vy;
      r = r + c[2][i] * vz;
      r = r + c[3][i] * vw;
      out[i] = r;
    }
  }
  Mat4 transpose() const {
    Mat4 t;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) t.c[i][j] = c[j][i];
    return t;
  }
};

// ---------------------------------------------------------------- RNG
struct Rng {
  virtual ~Rng() {}
  virtual float f32() = 0;                    // rng.random::<f32>()
  virtual Vec3 in_unit_sphere() = 0;          // Vec3::random_in_unit_sphere (vec3.rs:54-61)
  // Vec3::random_in_unit_sphere(rng).normalized() as used by Lambertian / Plastic (material.rs:55).  The ChaCha
  // stream does literally that; the counter-based streams map two uniforms straight onto the sphere (same
  // distribution, no wasted radius draw).
  virtual Vec3 unit_vector() { return in_unit_sphere().normalized(); }
  virtual void set_bounce(uint32_t) {}        // Philox mode only
};

// rand_chacha 0.9.0 ChaCha block.  words 12,13 = 64-bit block counter, 14,15 = 64-bit stream id.
inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  for (int i = 0; i < 8; i++) s[4 + i] = key[i];
  s[12] = (uint32_t)counter;
  s[13] = (uint32_t)(counter >> 32);
  s[14] = (uint32_t)stream;
  s[15] = (uint32_t)(stream >> 32);
  uint32_t x[16];
  std::memcpy(x, s, sizeof(x));
#define QR(a, b, c, d)                 \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); \
  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  \
  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
  for (int i = 0; i < rounds; i += 2) {
    QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
    QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
  }
#undef QR
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

// rand 0.9.1 StdRng = ChaCha12Rng; seed_from_u64 expands the u64 with PCG32 (rand_core 0.9 SeedableRng).
struct ChaChaRng : Rng {
  uint32_t key[8];
  uint64_t counter = 0;
  uint32_t buf[64];
  int idx = 64;
  explicit ChaChaRng(uint64_t seed) {
    uint64_t state = seed;
    for (int i = 0; i < 8; i++) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
      uint32_t rot = (uint32_t)(state >> 59);
      key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
  }
  uint32_t next_u32() {
    if (idx >= 64) {  // rand_chacha refills four blocks at a time
      for (int b = 0; b < 4; b++) chacha_block(key, counter + b, 0, 12, buf + 16 * b);
      counter += 4;
      idx = 0;
    }
    return buf[idx++];
  }
  float f32() override {  // rand: StandardUniform for f32 = 24 high bits * 2^-24
    return (float)(next_u32() >> 8) * (1.0f / 16777216.0f);
  }
  float range(float low, float high) {  // rand 0.9 UniformFloat::sample_single
    float scale = high - low;
    for (;;) {
      uint32_t bits = (next_u32() >> 9) | 0x3f800000u;  // [1,2)
      float v12;
      std::memcpy(&v12, &bits, 4);
      float v01 = v12 - 1.0f;
      float res = v01 * scale + low;
      if (res < high) return res;
    }
  }
  Vec3 in_unit_sphere() override {  // vec3.rs:46-61 (x, y, z drawn in that order)
    for (;;) {
      float x = range(-1.0f, 1.0f);
      float y = range(-1.0f, 1.0f);
      float z = range(-1.0f, 1.0f);
      Vec3 p(x, y, z);
      if (p.length_squared() < 1.0f) return p;
    }
  }
};

void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  for (int i = 0; i < 4; i++) out[i] = c[i];
}

inline float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// Uniform direction on S^2 from two uniforms (replaces normalised rejection sampling; same distribution).
inline Vec3 sphere_point(float u0, float u1) {
  float z = 1.0f - 2.0f * u0;
  float r = std::sqrt(rmax(0.0f, 1.0f - z * z));
  float phi = 2.0f * PI * u1;
  return Vec3(r * std::cos(phi), r * std::sin(phi), z);
}

// The GPU's stream: counter = (pixel, sample, bounce, block), key = seed.  Uniform k of a bounce is
// word k%4 of block k/4.  in_unit_sphere() = uniform ball point from the NEXT three uniforms.
struct PhiloxRng : Rng {
  uint32_t key[2];
  uint32_t pixel, sample, bounce = 0, dim = 0;
  uint32_t block[4];
  uint32_t have_block = 0xffffffffu;
  PhiloxRng(uint64_t seed, uint32_t px, uint32_t s) : pixel(px), sample(s) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
  }
  void set_bounce(uint32_t b) override {
    bounce = b;
    dim = 0;
    have_block = 0xffffffffu;
  }
  float f32() override {
    uint32_t blk = dim >> 2;
    if (blk != have_block) {
      uint32_t ctr[4] = {pixel, sample, bounce, blk};
      philox4x32_10(ctr, key, block);
      have_block = blk;
    }
    return u32_to_unit(block[(dim++) & 3]);
  }
  Vec3 in_unit_sphere() override {
    float u0 = f32(), u1 = f32(), u2 = f32();
    Vec3 d = sphere_point(u0, u1);
    return d * std::cbrt(u2);
  }
  Vec3 unit_vector() override {
    float u0 = f32(), u1 = f32();
    return sphere_point(u0, u1).normalized();
  }
};

// Fixed uniforms (orc_scatter hook).
struct ArrayRng : Rng {
  const float *u;
  int i = 0;
  explicit ArrayRng(const float *u_) : u(u_) {}
  float f32() override { return u[(i++) & 3]; }
  Vec3 in_unit_sphere() override {
    float u0 = f32(), u1 = f32(), u2 = f32();
    return sphere_point(u0, u1) * std::cbrt(u2);
  }
  Vec3 unit_vector() override {
    float u0 = f32(), u1 = f32();
    return sphere_point(u0, u1).normalized();
  }
};

// ---------------------------------------------------------------- hittable.rs
struct HitRecord {  // hittable.rs:10-16 (+ ids for parity reporting)
  Vec3 position, normal;
  float t = 0;
  int material = -1;
  bool front_face = false;
  int triangle = -1;
  void set_face_normal(const Ray &ray, Vec3 outward) {  // hittable.rs:19-26
    front_face = ray.direction.dot(outward) < 0.0f;
    normal = front_face ? outward : -outward;
  }
};

struct Hittable {
  virtual ~Hittable() {}
  virtual bool hit(const Ray &ray, float t_min, float t_max, HitRecord &rec) const = 0;
};

// ---------------------------------------------------------------- objects/sphere.rs:15-53
struct Sphere : Hittable {
  Vec3 center;
  float radius;
  int material;
  bool hit(const Ray &ray, float t_min, float t_max, HitRecord &rec) const override {
    Vec3 oc = ray.origin - center;
    float a = ray.direction.dot(ray.direction);
    float half_b = oc.dot(ray.direction);
    float c = oc.dot(oc) - radius * radius;
    float discriminant = half_b * half_b - a * c;
    if (discriminant < 0.0f) return false;
    float sqrtd = std::sqrt(discriminant);
    float root = (-half_b - sqrtd) / a;
    if (root <= t_min || root >= t_max) {
      root = (-half_b + sqrtd) / a;
      if (root <= t_min || root >= t_max) return false;
    }
    rec.t = root;
    rec.position = ray.at(root);
    Vec3 outward = vdiv(rec.position - center, radius);
    rec.front_face = ray.direction.dot(outward) < 0.0f;
    rec.normal = rec.front_face ? outward : -outward;
    rec.material = material;
    rec.triangle = -1;
    return true;
  }
};

// ---------------------------------------------------------------- objects/plane.rs:26-56
struct Plane : Hittable {
  Vec3 p1, normal;
  int material;
  bool hit(const Ray &ray, float t_min, float t_max, HitRecord &rec) const override {
    float denom = normal.dot(ray.direction);
    if (std::fabs(denom) < EPSILON) return false;
    float t = normal.dot(p1 - ray.origin) / denom;
    if (t <= t_min || t >= t_max) return false;
    rec.t = t;
    rec.position = ray.at(t);
    rec.front_face = ray.direction.dot(normal) < 0.0f;
    rec.normal = rec.front_face ? normal : -normal;
    rec.material = material;
    rec.triangle = -1;
    return true;
  }
};

// ---------------------------------------------------------------- tungsten/objects/quad.rs:83-132
struct Quad : Hittable {
  Vec3 base, edge0, edge1, normal;
  float d, inv_edge0_len_sq, inv_edge1_len_sq;
  int material;
  bool hit(const Ray &ray, float t_min, float t_max, HitRecord &rec) const override {
    float denom = normal.dot(ray.direction);
    if (std::fabs(denom) < EPSILON) return false;
    float t = (d - normal.dot(ray.origin)) / denom;
    if (t <= t_min || t >= t_max) return false;
    Vec3 hit_pos = ray.at(t);
    Vec3 v = hit_pos - base;
    float l0 = v.dot(edge0) * inv_edge0_len_sq;
    float l1 = v.dot(edge1) * inv_edge1_len_sq;
    const float lo = -EPSILON, hi = 1.0f + EPSILON;  // (-EPSILON..=1.0+EPSILON).contains()
    if (!((lo <= l0 && l0 <= hi) && (lo <= l1 && l1 <= hi))) return false;
    rec.t = t;
    rec.position = hit_pos;
    rec.front_face = ray.direction.dot(normal) < 0.0f;
    rec.normal = rec.front_face ? normal : -normal;
    rec.material = material;
    rec.triangle = -1;
    return true;
  }
};

// f32::signum: 1.0 for +0.0 and positives, -1.0 for -0.0 and negatives, NaN for NaN.
inline float rsignum(float v) {
  if (std::isnan(v)) return v;
  return std::signbit(v) ? -1.0f : 1.0f;
}

// ---------------------------------------------------------------- objects/cube.rs:59-158
struct Cube : Hittable {
  Mat4 o2w, w2o;
  int material;
  bool hit(const Ray &ray, float t_min, float t_max, HitRecord &rec) const override {
    float oh[4], dh[4];
    w2o.mul(ray.origin.x, ray.origin.y, ray.origin.z, 1.0f, oh);           // cube.rs:63-69
    w2o.mul(ray.direction.x, ray.direction.y, ray.direction.z, 0.0f, dh);  // cube.rs:70-76
    float o[3] = {oh[0], oh[1], oh[2]}, dd[3] = {dh[0], dh[1], dh[2]};
    float te[3], tx[3];
    for (int i = 0; i < 3; i++) {  // cube.rs:85-90
      float inv = 1.0f / dd[i];
      float t1 = (-0.5f - o[i]) * inv;
      float t2 = (0.5f - o[i]) * inv;
      te[i] = rmin(t1, t2);
      tx[i] = rmax(t1, t2);
    }
    float t_enter = rmax(te[0], rmax(te[1], te[2]));  // cube.rs:92
    float t_exit = rmin(tx[0], rmin(tx[1], tx[2]));   // cube.rs:93
    if (t_exit < t_enter || t_exit <= 0.0f) return false;  // cube.rs:95
    float t_hit_obj = t_enter > 0.0f ? t_enter : t_exit;    // cube.rs:99
    if (t_hit_obj >= t_max || t_hit_obj <= t_min || t_hit_obj < EPSILON) return false;  // cube.rs:101
    float p[3];
    for (int i = 0; i < 3; i++) p[i] = o[i] + dd[i] * t_hit_obj;  // cube.rs:105
    float n[3] = {0, 0, 0};
    float ax = std::fabs(p[0]), ay = std::fabs(p[1]), az = std::fabs(p[2]);
    const float tol = 1e-4f;
    if (std::fabs(ax - 0.5f) < tol) n[0] = rsignum(p[0]);  // cube.rs:111-123
    else if (std::fabs(ay - 0.5f) < tol) n[1] = rsignum(p[1]);
    else if (std::fabs(az - 0.5f) < tol) n[2] = rsignum(p[2]);
    else if (ax > ay && ax > az) n[0] = rsignum(p[0]);
    else if (ay > az) n[1] = rsignum(p[1]);
    else n[2] = rsignum(p[2]);
    {  // glam normalize_or_zero (cube.rs:124)
      float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      float rcp = 1.0f / len;
      if (std::isfinite(rcp) && rcp > 0.0f) {
        n[0] *= rcp; n[1] *= rcp; n[2] *= rcp;
      } else {
        n[0] = n[1] = n[2] = 0.0f;
      }
    }
    float pw[4], nw[4];
    o2w.mul(p[0], p[1], p[2], 1.0f, pw);                    // cube.rs:126-129
    w2o.transpose().mul(n[0], n[1], n[2], 0.0f, nw);        // cube.rs:131-134
    Vec3 position_world(pw[0], pw[1], pw[2]);
    Vec3 normal_world = Vec3(nw[0], nw[1], nw[2]).normalized();
    Vec3 p_minus_o = position_world - ray.origin;
    if (p_minus_o.dot(ray.direction) < 0.0f) return false;  // cube.rs:144-147
    float t_world = (position_world - ray.origin).dot(ray.direction);
    if (t_world < t_min || t_world > t_max) return false;   // cube.rs:150 (closed interval)
    rec.t = t_world;
    rec.position = position_world;
    rec.material = material;
    rec.triangle = -1;
    rec.set_face_normal(ray, normal_world);  // cube.rs:155
    return true;
  }
};

// ---------------------------------------------------------------- mesh/triangle.rs, acceleration/{aabb,bvh}.rs
struct Triangle {
  Vec3 v0, v1, v2, normal;
};

struct Aabb {
  Vec3 min, max;
  Aabb() {  // aabb.rs:11-16
    float inf = std::numeric_limits<float>::infinity();
    min = Vec3(inf, inf, inf);
    max = Vec3(-inf, -inf, -inf);
  }
  void add_point(Vec3 p) {  // aabb.rs:18-25
    min.x = rmin(min.x, p.x); min.y = rmin(min.y, p.y); min.z = rmin(min.z, p.z);
    max.x = rmax(max.x, p.x); max.y = rmax(max.y, p.y); max.z = rmax(max.z, p.z);
  }
  bool intersect(const Ray &ray, float t_min, float t_max) const {  // aabb.rs:27-45
    for (int axis = 0; axis < 3; axis++) {
      float inv_d = 1.0f / ray.direction[axis];
      float t0 = (min[axis] - ray.origin[axis]) * inv_d;
      float t1 = (max[axis] - ray.origin[axis]) * inv_d;
      if (inv_d < 0.0f) std::swap(t0, t1);
      t_min = rmax(t_min, t0);
      t_max = rmin(t_max, t1);
      if (t_max <= t_min) return false;
    }
    return true;
  }
};

struct BVHNode {  // bvh.rs:7-12
  Aabb bounds;
  std::unique_ptr<BVHNode> left, right;
  std::vector<size_t> triangle_indices;
};

// bvh.rs:15-76.  `indices` is sorted in place, so after the build the index array is the DFS leaf order.
// Rust's sort_unstable_by tie order depends on the std version (pdqsort up to 1.80, ipnsort since; neither can be run or
// pinned here): the oracle uses a STABLE sort by default (divergence class T6).  PTC_REF_TIE selects another order of
// equal keys — `reverse` (reversed input order) or `random:<seed>` (hash of seed and triangle index) — so that the
// sensitivity of the flat-node set to that unspecified order can be measured (tools/tie_order_study.py).
struct TieOrder {
  int kind = 0;  // 0 stable, 1 reverse, 2 random
  uint64_t seed = 0;
  static TieOrder from_env() {
    TieOrder t;
    const char *e = getenv("PTC_REF_TIE");
    if (!e || !*e || !strcmp(e, "stable")) return t;
    if (!strcmp(e, "reverse")) t.kind = 1;
    else if (!strncmp(e, "random:", 7)) t.kind = 2, t.seed = strtoull(e + 7, nullptr, 10);
    return t;
  }
  static uint64_t mix(uint64_t seed, uint64_t i) {  // splitmix64
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
};
std::unique_ptr<BVHNode> bvh_build(const std::vector<Triangle> &tris, size_t *idx, size_t n, size_t depth, const TieOrder &tie = TieOrder()) {
  auto node = std::make_unique<BVHNode>();
  for (size_t i = 0; i < n; i++) {
    const Triangle &t = tris[idx[i]];
    node->bounds.add_point(t.v0);
    node->bounds.add_point(t.v1);
    node->bounds.add_point(t.v2);
  }
  const size_t MAX_DEPTH = 25, MIN_TRIANGLES_PER_LEAF = 4;
  if (n <= MIN_TRIANGLES_PER_LEAF || depth >= MAX_DEPTH) {
    node->triangle_indices.assign(idx, idx + n);
    return node;
  }
  Vec3 extent = node->bounds.max - node->bounds.min;
  int axis = (extent.x > extent.y && extent.x > extent.z) ? 0 : (extent.y > extent.z ? 1 : 2);
  auto centroid = [&](size_t i) {
    const Triangle &t = tris[i];
    return ((t.v0 + t.v1 + t.v2) * (1.0f / 3.0f))[axis];
  };
  if (tie.kind == 0) {
    std::stable_sort(idx, idx + n, [&](size_t a, size_t b) {
      float va = centroid(a), vb = centroid(b);
      return va < vb;  // partial_cmp(..).unwrap_or(Equal): NaN compares equal
    });
  } else {
    std::vector<std::pair<uint64_t, size_t>> sec(n);  // (secondary key realising the tie order, triangle)
    for (size_t i = 0; i < n; i++) sec[i] = {tie.kind == 1 ? (uint64_t)(n - 1 - i) : TieOrder::mix(tie.seed, (uint64_t)idx[i]), idx[i]};
    std::stable_sort(sec.begin(), sec.end(), [&](const std::pair<uint64_t, size_t> &a, const std::pair<uint64_t, size_t> &b) {
      float va = centroid(a.second), vb = centroid(b.second);
      if (va < vb) return true;
      if (vb < va) return false;
      return a.first < b.first;
    });
    for (size_t i = 0; i < n; i++) idx[i] = sec[i].second;
  }
  size_t mid = n / 2;
  if (mid == 0 || mid == n) {
    node->triangle_indices.assign(idx, idx + n);
    return node;
  }
  node->left = bvh_build(tris, idx, mid, depth + 1, tie);
  node->right = bvh_build(tris, idx + mid, n - mid, depth + 1, tie);
  return node;
}

// bvh.rs:78-170
bool bvh_intersect(const BVHNode &node, const Ray &ray, const std::vector<Triangle> &tris, float t_min, float t_max,
                   HitRecord &out) {
  if (!node.bounds.intersect(ray, t_min, t_max)) return false;
  if (!node.left) {
    bool any = false;
    for (size_t idx : node.triangle_indices) {
      const Triangle &tr = tris[idx];
      Vec3 edge1 = tr.v1 - tr.v0;
      Vec3 edge2 = tr.v2 - tr.v0;
      Vec3 h = ray.direction.cross(edge2);
      float a = edge1.dot(h);
      if (std::fabs(a) < EPSILON) continue;
      float f = 1.0f / a;
      Vec3 s = ray.origin - tr.v0;
      float u = f * s.dot(h);
      if (!(0.0f <= u && u <= 1.0f)) continue;
      Vec3 q = s.cross(edge1);
      float v = f * ray.direction.dot(q);
      if (v < 0.0f || u + v > 1.0f) continue;
      float t = f * edge2.dot(q);
      if (t > t_min && t < t_max) {
        out.t = t;
        out.position = ray.at(t);
        out.front_face = ray.direction.dot(tr.normal) < 0.0f;
        out.normal = out.front_face ? tr.normal : -tr.normal;
        out.triangle = (int)idx;
        t_max = t;
        any = true;
      }
    }
    return any;
  }
  HitRecord l, r;
  bool hl = bvh_intersect(*node.left, ray, tris, t_min, t_max, l);
  if (hl) t_max = l.t;
  bool hr = bvh_intersect(*node.right, ray, tris, t_min, t_max, r);
  if (hl && hr) {
    out = (l.t < r.t) ? l : r;
    return true;
  }
  if (hl) { out = l; return true; }
  if (hr) { out = r; return true; }
  return false;
}

// ---------------------------------------------------------------- mesh/mesh_object.rs:262-329
struct Mesh : Hittable {
  std::vector<Triangle> triangles;
  std::unique_ptr<BVHNode> bvh;
  std::vector<size_t> order;  // index array after the build = DFS leaf order
  Mat4 o2w, w2o;
  int material;
  bool hit(const Ray &ray_world, float t_min_world, float t_max_world, HitRecord &rec) const override {
    float oh[4], dh[4];
    w2o.mul(ray_world.origin.x, ray_world.origin.y, ray_world.origin.z, 1.0f, oh);
    w2o.mul(ray_world.direction.x, ray_world.direction.y, ray_world.direction.z, 0.0f, dh);
    Vec3 ray_origin_obj(oh[0], oh[1], oh[2]);
    Vec3 ray_direction_obj(dh[0], dh[1], dh[2]);
    Ray ray_obj(ray_origin_obj, ray_direction_obj.normalized());  // normalised twice (mesh_object.rs:287 + ray.rs:15)
    HitRecord h;
    if (!bvh_intersect(*bvh, ray_obj, triangles, t_min_world, t_max_world, h)) return false;
    float pw[4], nw[4];
    o2w.mul(h.position.x, h.position.y, h.position.z, 1.0f, pw);
    w2o.transpose().mul(h.normal.x, h.normal.y, h.normal.z, 0.0f, nw);
    Vec3 pos_world(pw[0], pw[1], pw[2]);
    Vec3 normal_world = Vec3(nw[0], nw[1], nw[2]).normalized();
    float ray_dir_obj_length = ray_direction_obj.length();
    float ray_dir_world_length = ray_world.direction.length();
    float t_world = h.t * ray_dir_obj_length / ray_dir_world_length;  // mesh_object.rs:312-314
    if (t_world < t_min_world || t_world > t_max_world) return false; // closed (mesh_object.rs:316)
    rec = h;
    rec.position = pos_world;
    rec.t = t_world;
    rec.material = material;
    rec.set_face_normal(ray_world, normal_world);
    return true;
  }
};

// ---------------------------------------------------------------- material.rs / tungsten/materials.rs
Vec3 reflect_checked(Vec3 v_in, Vec3 n) {  // material.rs:194-206, tungsten/materials.rs:292-304
  float nan = std::numeric_limits<float>::quiet_NaN();
  if (v_in.has_nan()) return Vec3(nan, nan, nan);
  if (n.has_nan() || n.is_zero()) return Vec3(nan, nan, nan);
  return v_in - n * 2.0f * v_in.dot(n);
}

inline float powi5(float x) {  // f32::powi(5): x * (x^2)^2
  float x2 = x * x;
  float x4 = x2 * x2;
  return x * x4;
}

float schlick_reflectance(float cosine, float ref_idx_ratio) {  // material.rs:221-227
  float r0 = (1.0f - ref_idx_ratio) / (1.0f + ref_idx_ratio);
  r0 = r0 * r0;
  return r0 + (1.0f - r0) * powi5(1.0f - cosine);
}

float schlick_plastic(float cosine, float ref_idx) {  // tungsten/materials.rs:23-27
  float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
  float r0_sq = r0 * r0;
  return r0_sq + (1.0f - r0_sq) * powi5(1.0f - cosine);
}

bool refract(Vec3 uv, Vec3 n, float etai_over_etat, Vec3 &out) {  // material.rs:208-219
  float cos_theta = rmin((-uv).dot(n), 1.0f);
  Vec3 r_out_perp = (uv + n * cos_theta) * etai_over_etat;
  float r_out_parallel_squared = 1.0f - r_out_perp.length_squared();
  if (r_out_parallel_squared < 0.0f) return false;
  Vec3 r_out_parallel = n * (-std::sqrt(r_out_parallel_squared));
  out = r_out_perp + r_out_parallel;
  return true;
}

Color checker_value(const orc_material &m, Vec3 p) {  // tungsten/materials.rs:89-99
  // `as i32` saturates; wrapping add is what release-mode Rust does on overflow.
  auto to_i32 = [](float f) -> int32_t {
    if (std::isnan(f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
  };
  int32_t xc = to_i32(std::floor(p.x * m.inv_scale));
  int32_t yc = to_i32(std::floor(p.y * m.inv_scale));
  int32_t zc = to_i32(std::floor(p.z * m.inv_scale));
  int32_t s = (int32_t)((uint32_t)xc + (uint32_t)yc + (uint32_t)zc);
  if (s % 2 == 0) return Color(m.albedo[0], m.albedo[1], m.albedo[2]);
  return Color(m.off_color[0], m.off_color[1], m.off_color[2]);
}

Color fresnel_conductor(float cos_theta, Color eta, Color k) {  // tungsten/materials.rs:184-202
  cos_theta = cos_theta < 0.0f ? 0.0f : (cos_theta > 1.0f ? 1.0f : cos_theta);
  Color cos2 = Color::splat(cos_theta * cos_theta);
  Color sin2 = Color::splat(1.0f) - cos2;
  Color eta2 = eta * eta;
  Color k2 = k * k;
  Color t0 = eta2 - k2 - sin2;
  Color a2plusb2 = (t0 * t0 + Color::splat(4.0f) * eta2 * k2).sqrt();
  Color t1 = a2plusb2 + cos2;
  Color a = (a2plusb2 + t0) * Color::splat(0.5f);
  a = a.sqrt();
  Color t2 = Color::splat(2.0f * cos_theta) * a;
  Color rs = (t1 - t2) / (t1 + t2);
  Color t3 = cos2 * a2plusb2 + sin2 * sin2;
  Color t4 = t2;
  Color rp = rs * ((t3 - t4) / (t3 + t4));
  return (rs + rp) * Color::splat(0.5f);
}

float ggx_g1(float n_dot_x, float roughness) {  // tungsten/materials.rs:205-216
  if (n_dot_x <= 0.0f) return 0.0f;
  float a = roughness * roughness;
  float k = a / 2.0f;
  float denom = n_dot_x * (1.0f - k) + k;
  if (denom < EPSILON) return 1.0f;
  return n_dot_x / denom;
}
float ggx_g(float roughness, float n_dot_v, float n_dot_l) {  // :218-221
  return ggx_g1(n_dot_v, roughness) * ggx_g1(n_dot_l, roughness);
}
float beckmann_g(float roughness, float n_dot_v, float n_dot_l) {  // :223-234
  float a = roughness;
  auto lambda = [a](float x) {
    float t = 1.0f / (a * x);
    if (t < 1.6f) return (1.0f - 1.259f * t + 0.396f * t * t) / (3.535f * t + 2.181f * t * t);
    return 0.0f;
  };
  return 1.0f / (1.0f + lambda(n_dot_v) + lambda(n_dot_l));
}

Vec3 sample_half_vector(Vec3 normal, float roughness, int distribution, Rng &rng) {  // :236-290
  float nan = std::numeric_limits<float>::quiet_NaN();
  if (normal.has_nan() || normal.is_zero()) return Vec3(nan, nan, nan);
  float u1 = rmax(rng.f32(), 1e-6f);
  float u2 = rng.f32();
  float theta_arg;
  if (distribution == ORC_DIST_GGX) {
    float a = roughness * roughness;
    theta_arg = a * a * (-std::log(u1)) / (1.0f - u1);
  } else {
    theta_arg = -(roughness * roughness * std::log(u1));
  }
  if (std::isnan(theta_arg) || std::isinf(theta_arg) || theta_arg < 0.0f) return to_world(Vec3(0, 0, 1), normal);
  float theta = std::atan(std::sqrt(theta_arg));
  float phi = 2.0f * PI * u2;
  float sin_theta = std::sin(theta), cos_theta = std::cos(theta);
  Vec3 h_local(sin_theta * std::cos(phi), sin_theta * std::sin(phi), cos_theta);
  if (h_local.has_nan()) return to_world(Vec3(0, 0, 1), normal);
  return to_world(h_local, normal);
}

Vec3 lambert_direction(const HitRecord &hit, Rng &rng) {  // material.rs:54-59, tungsten/materials.rs:55-59
  Vec3 d = hit.normal + rng.unit_vector();
  if (d.near_zero()) d = hit.normal;
  return d;
}

Color mat_emitted(const orc_material &m) {  // material.rs:18-20, 188-190
  if (m.type == ORC_MAT_EMISSIVE) return Color(m.albedo[0], m.albedo[1], m.albedo[2]);
  return Color();
}

bool mat_scatter(const orc_material &m, const Ray &ray_in, const HitRecord &hit, Rng &rng, Ray &scattered,
                 Color &attenuation) {
  switch (m.type) {
    case ORC_MAT_LAMBERT:
    case ORC_MAT_LAMBERT_CHECKER: {  // material.rs:48-70
      Vec3 dir = lambert_direction(hit, rng);
      Vec3 origin = hit.position + hit.normal * EPSILON;
      scattered = Ray(origin, dir.normalized());
      attenuation = m.type == ORC_MAT_LAMBERT ? Color(m.albedo[0], m.albedo[1], m.albedo[2]) : checker_value(m, hit.position);
      return true;
    }
    case ORC_MAT_METAL: {  // material.rs:88-109
      Vec3 reflected = reflect_checked(ray_in.direction.normalized(), hit.normal);
      Vec3 fuzzed = m.fuzz > 0.0f ? reflected + rng.in_unit_sphere() * m.fuzz : reflected;
      if (fuzzed.dot(hit.normal) > 0.0f) {
        Vec3 origin = hit.position + hit.normal * EPSILON;
        scattered = Ray(origin, fuzzed.normalized());
        attenuation = Color(m.albedo[0], m.albedo[1], m.albedo[2]);
        return true;
      }
      return false;
    }
    case ORC_MAT_DIELECTRIC: {  // material.rs:123-162
      float ratio = hit.front_face ? 1.0f / m.ior : m.ior / 1.0f;
      Vec3 unit = ray_in.direction.normalized();
      float cos_theta = rmin((-unit).dot(hit.normal), 1.0f);
      float sin_theta_squared = 1.0f - cos_theta * cos_theta;
      bool cannot_refract = ratio * ratio * sin_theta_squared > 1.0f;
      float reflectance = schlick_reflectance(cos_theta, 1.0f / ratio);
      Vec3 dir;
      if (cannot_refract || reflectance > rng.f32()) {  // short-circuit: no draw on TIR
        dir = reflect_checked(unit, hit.normal);
      } else {
        Vec3 refr;
        dir = refract(unit, hit.normal, ratio, refr) ? refr : reflect_checked(unit, hit.normal);
      }
      Vec3 origin = dir.dot(hit.normal) > 0.0f ? hit.position + hit.normal * EPSILON : hit.position - hit.normal * EPSILON;
      scattered = Ray(origin, dir.normalized());
      attenuation = Color(1, 1, 1);
      return true;
    }
    case ORC_MAT_EMISSIVE:  // material.rs:179-186
    case ORC_MAT_NULL:      // material.rs:239-247
      return false;
    case ORC_MAT_PLASTIC: {  // tungsten/materials.rs:30-65
      float dn = ray_in.direction.dot(hit.normal);
      float cosine = dn > 0.0f ? m.ior * dn / ray_in.direction.length() : -dn / ray_in.direction.length();
      float reflect_prob = schlick_plastic(cosine, m.ior);
      if (rng.f32() < reflect_prob) {
        Vec3 reflected = ray_in.direction.reflect(hit.normal).normalized();
        Vec3 origin = hit.position + hit.normal * EPSILON;
        scattered = Ray(origin, reflected);
        attenuation = Color(0.9f, 0.9f, 0.9f);
      } else {
        Vec3 dir = lambert_direction(hit, rng);
        Vec3 origin = hit.position + hit.normal * EPSILON;
        scattered = Ray(origin, dir.normalized());
        attenuation = Color(m.albedo[0], m.albedo[1], m.albedo[2]);
      }
      return true;
    }
    case ORC_MAT_ROUGH_CONDUCTOR: {  // tungsten/materials.rs:307-376
      if (ray_in.direction.has_nan()) return false;
      if (hit.normal.has_nan() || hit.normal.is_zero()) return false;
      Vec3 n = hit.normal;
      Vec3 v = -ray_in.direction.normalized();
      if (v.has_nan()) return false;
      Color eta(m.eta[0], m.eta[1], m.eta[2]), k(m.k[0], m.k[1], m.k[2]);
      float rough = m.roughness;
      Vec3 h = sample_half_vector(n, rough, m.distribution, rng);
      if (h.has_nan()) return false;
      Vec3 l = reflect_checked(-v, h);
      if (l.has_nan()) return false;
      if (l.dot(n) <= 0.0f) return false;
      float n_dot_l = rmax(n.dot(l), 0.0f);
      float n_dot_v = rmax(n.dot(v), 0.0f);
      float n_dot_h = rmax(n.dot(h), 0.0f);
      float v_dot_h = rmax(v.dot(h), 0.0f);
      float g = m.distribution == ORC_DIST_GGX ? ggx_g(rough, n_dot_v, n_dot_l) : beckmann_g(rough, n_dot_v, n_dot_l);
      Color f = fresnel_conductor(v_dot_h, eta, k);
      Color num = f * g * v_dot_h;
      float den = n_dot_v * n_dot_h + EPSILON;
      Color color = den > EPSILON ? Color(m.albedo[0], m.albedo[1], m.albedo[2]) * (num / den) : Color();
      Vec3 origin = hit.position + n * EPSILON;
      scattered = Ray(origin, l.normalized());
      attenuation = color;
      return true;
    }
  }
  return false;
}

}  // namespace

// ---------------------------------------------------------------- scene.rs + hittable.rs:28-57
struct orc_scene {
  std::vector<orc_material> materials;
  std::vector<std::unique_ptr<Hittable>> objects;
  std::vector<float> sky;
  int sky_w = 0, sky_h = 0;

  bool hit(const Ray &ray, float t_min, float t_max, HitRecord &rec, int &object) const {  // hittable.rs:46-57
    float closest = t_max;
    bool any = false;
    for (size_t i = 0; i < objects.size(); i++) {
      HitRecord tmp;
      if (objects[i]->hit(ray, t_min, closest, tmp)) {
        closest = tmp.t;
        rec = tmp;
        object = (int)i;
        any = true;
      }
    }
    return any;
  }
};

namespace {

Color sky_color(const orc_scene &sc, const Ray &ray) {  // renderer.rs:38-63
  if (sc.sky_w > 0) {
    Vec3 dir = ray.direction.normalized();
    float theta = std::acos(dir.y);
    float phi = std::atan2(dir.z, dir.x) + PI;
    float u = phi / (2.0f * PI);
    float v = theta / PI;
    // `as u32` saturates (NaN -> 0)
    auto to_u32 = [](float f) -> uint32_t {
      if (!(f > 0.0f)) return 0u;
      if (f >= 4294967296.0f) return 0xffffffffu;
      return (uint32_t)f;
    };
    uint32_t xp = to_u32(rmax(u * (float)(sc.sky_w - 1), 0.0f));
    uint32_t yp = to_u32(rmax(v * (float)(sc.sky_h - 1), 0.0f));
    xp = std::min(xp, (uint32_t)(sc.sky_w - 1));
    yp = std::min(yp, (uint32_t)(sc.sky_h - 1));
    const float *px = &sc.sky[((size_t)yp * sc.sky_w + xp) * 3];
    return Color(px[0], px[1], px[2]);
  }
  return Color(0.5f, 0.5f, 0.5f);  // Color::GRAY
}

// renderer.rs:19-65.  `bounce` = number of segments already traced (0 for the camera ray); it only keys the
// Philox stream and is not part of the reference's signature.
Color trace_ray(const Ray &ray_in, const orc_scene &sc, size_t depth, Rng &rng, uint32_t bounce, uint64_t &rays) {
  if (depth == 0) return Color();
  rays++;
  HitRecord hit;
  int object = -1;
  if (sc.hit(ray_in, EPSILON, std::numeric_limits<float>::infinity(), hit, object)) {
    const orc_material &m = sc.materials[hit.material];
    Color emitted = mat_emitted(m);
    Ray scattered;
    Color attenuation;
    rng.set_bounce(bounce);
    if (mat_scatter(m, ray_in, hit, rng, scattered, attenuation)) {
      Color c = trace_ray(scattered, sc, depth - 1, rng, bounce + 1, rays);
      return emitted + attenuation * c;
    }
    return emitted;
  }
  return sky_color(sc, ray_in);
}

Ray camera_get_ray(const orc_camera &c, float u, float v) {  // camera.rs:33-42
  float ndc_x = 2.0f * u - 1.0f;
  float ndc_y = 1.0f - 2.0f * v;
  Vec3 right(c.right[0], c.right[1], c.right[2]), up(c.true_up[0], c.true_up[1], c.true_up[2]);
  Vec3 fwd(c.forward[0], c.forward[1], c.forward[2]);
  Vec3 offset = right * (ndc_x * c.half_width) + up * (ndc_y * c.half_height);
  Vec3 dir = (fwd + offset).normalized();
  return Ray(Vec3(c.position[0], c.position[1], c.position[2]), dir);
}

Mat4 mat_from(const float m[16]) {
  Mat4 r;
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 4; i++) r.c[c][i] = m[c * 4 + i];
  return r;
}

void bvh_walk(const BVHNode &n, int depth, bool dead, int64_t &nodes, int64_t &leaves, int &maxd, uint8_t *dead_out) {
  nodes++;
  maxd = std::max(maxd, depth);
  Vec3 e = n.bounds.max - n.bounds.min;
  bool flat = (n.bounds.min.x == n.bounds.max.x) || (n.bounds.min.y == n.bounds.max.y) || (n.bounds.min.z == n.bounds.max.z);
  (void)e;
  dead = dead || flat;
  if (!n.left) {
    leaves++;
    if (dead_out)
      for (size_t i : n.triangle_indices) dead_out[i] = dead ? 1 : 0;
    return;
  }
  bvh_walk(*n.left, depth + 1, dead, nodes, leaves, maxd, dead_out);
  bvh_walk(*n.right, depth + 1, dead, nodes, leaves, maxd, dead_out);
}

}  // namespace

// ---------------------------------------------------------------- C API
extern "C" {

orc_scene *orc_scene_create(void) { return new orc_scene(); }
void orc_scene_destroy(orc_scene *s) { delete s; }

int orc_scene_add_material(orc_scene *s, const orc_material *m) {
  s->materials.push_back(*m);
  return (int)s->materials.size() - 1;
}

int orc_scene_add_sphere(orc_scene *s, const float center[3], float radius, int material) {
  auto o = std::make_unique<Sphere>();
  o->center = Vec3(center[0], center[1], center[2]);
  o->radius = radius;
  o->material = material;
  s->objects.push_back(std::move(o));
  return (int)s->objects.size() - 1;
}

int orc_scene_add_plane(orc_scene *s, const float p1[3], const float normal[3], int material) {
  auto o = std::make_unique<Plane>();
  o->p1 = Vec3(p1[0], p1[1], p1[2]);
  o->normal = Vec3(normal[0], normal[1], normal[2]);  // Plane::new already normalised it (plane.rs:19)
  o->material = material;
  s->objects.push_back(std::move(o));
  return (int)s->objects.size() - 1;
}

int orc_scene_add_quad(orc_scene *s, const float base[3], const float e0[3], const float e1[3], const float normal[3],
                       float d, float inv0, float inv1, int material) {
  auto o = std::make_unique<Quad>();
  o->base = Vec3(base[0], base[1], base[2]);
  o->edge0 = Vec3(e0[0], e0[1], e0[2]);
  o->edge1 = Vec3(e1[0], e1[1], e1[2]);
  o->normal = Vec3(normal[0], normal[1], normal[2]);
  o->d = d;
  o->inv_edge0_len_sq = inv0;
  o->inv_edge1_len_sq = inv1;
  o->material = material;
  s->objects.push_back(std::move(o));
  return (int)s->objects.size() - 1;
}

int orc_scene_add_cube(orc_scene *s, const float o2w[16], const float w2o[16], int material) {
  auto o = std::make_unique<Cube>();
  o->o2w = mat_from(o2w);
  o->w2o = mat_from(w2o);
  o->material = material;
  s->objects.push_back(std::move(o));
  return (int)s->objects.size() - 1;
}

int orc_scene_add_mesh(orc_scene *s, const float *tris, int64_t n, const float o2w[16], const float w2o[16], int material) {
  if (n <= 0) return -1;  // mesh_object.rs:30-36
  auto o = std::make_unique<Mesh>();
  o->triangles.resize((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    const float *p = tris + i * 12;
    Triangle &t = o->triangles[(size_t)i];
    t.v0 = Vec3(p[0], p[1], p[2]);
    t.v1 = Vec3(p[3], p[4], p[5]);
    t.v2 = Vec3(p[6], p[7], p[8]);
    t.normal = Vec3(p[9], p[10], p[11]);
  }
  o->order.resize((size_t)n);
  for (size_t i = 0; i < (size_t)n; i++) o->order[i] = i;
  o->bvh = bvh_build(o->triangles, o->order.data(), (size_t)n, 0, TieOrder::from_env());  // mesh_object.rs:44-45
  o->o2w = mat_from(o2w);
  o->w2o = mat_from(w2o);
  o->material = material;
  s->objects.push_back(std::move(o));
  return (int)s->objects.size() - 1;
}

int orc_scene_set_sky_hdr(orc_scene *s, const float *rgb, int w, int h) {
  s->sky.assign(rgb, rgb + (size_t)w * h * 3);
  s->sky_w = w;
  s->sky_h = h;
  return 0;
}

int orc_intersect(const orc_scene *s, const float *origins, const float *dirs, int64_t n, float t_min, float t_max,
                  orc_hit *out) {
  unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    const int64_t chunk = 4096;
    for (;;) {
      int64_t b = next.fetch_add(chunk);
      if (b >= n) break;
      int64_t e = std::min(n, b + chunk);
      for (int64_t i = b; i < e; i++) {
        Ray r = Ray::raw(Vec3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]),
                         Vec3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]));
        HitRecord h;
        int obj = -1;
        orc_hit &o = out[i];
        if (s->hit(r, t_min, t_max, h, obj)) {
          o.object = obj;
          o.triangle = h.triangle;
          o.t = h.t;
          o.position[0] = h.position.x; o.position[1] = h.position.y; o.position[2] = h.position.z;
          o.normal[0] = h.normal.x; o.normal[1] = h.normal.y; o.normal[2] = h.normal.z;
          o.front_face = h.front_face ? 1 : 0;
          o.material = h.material;
        } else {
          std::memset(&o, 0, sizeof(o));
          o.object = -1;
          o.triangle = -1;
          o.material = -1;
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (unsigned i = 1; i < nt; i++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  return 0;
}

int orc_render(const orc_scene *s, const orc_camera *cam, const orc_settings *st, float *image, orc_stats *stats) {
  const int W = st->width, H = st->height;
  const int s_begin = (st->sample_begin == 0 && st->sample_end == 0) ? 0 : st->sample_begin;
  const int s_end = (st->sample_begin == 0 && st->sample_end == 0) ? st->spp : st->sample_end;
  unsigned nt = st->threads > 0 ? (unsigned)st->threads : std::max(1u, std::thread::hardware_concurrency());
  std::atomic<int> next_row{0};
  std::atomic<uint64_t> total_rays{0};
  const float inv_spp = 1.0f / (float)st->spp;  // renderer.rs:85
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&]() {
    uint64_t rays = 0;
    for (;;) {  // renderer.rs:87-106: rows are the parallel unit (par_chunks_mut(width)), scheduled dynamically
      int y = next_row.fetch_add(1);
      if (y >= H) break;
      std::unique_ptr<ChaChaRng> row_rng;
      if (st->rng_mode == ORC_RNG_CHACHA) row_rng = std::make_unique<ChaChaRng>((uint64_t)y);  // renderer.rs:91
      for (int x = 0; x < W; x++) {
        Color acc;
        for (int sidx = s_begin; sidx < s_end; sidx++) {
          Color c;
          if (st->rng_mode == ORC_RNG_CHACHA) {
            float u = ((float)x + row_rng->f32()) / (float)W;  // renderer.rs:96
            float v = ((float)y + row_rng->f32()) / (float)H;  // renderer.rs:97
            Ray ray = camera_get_ray(*cam, u, v);
            c = trace_ray(ray, *s, (size_t)st->max_depth, *row_rng, 0, rays);
          } else {
            PhiloxRng rng(st->seed, (uint32_t)(y * W + x), (uint32_t)sidx);
            rng.set_bounce(0xffffffffu);
            float ju = rng.f32(), jv = rng.f32();
            float u = ((float)x + ju) / (float)W;
            float v = ((float)y + jv) / (float)H;
            Ray ray = camera_get_ray(*cam, u, v);
            c = trace_ray(ray, *s, (size_t)st->max_depth, rng, 0, rays);
          }
          acc = acc + c;
        }
        Color px = acc * inv_spp;  // renderer.rs:103
        float *o = image + ((size_t)y * W + x) * 3;
        o[0] = px.r; o[1] = px.g; o[2] = px.b;
      }
    }
    total_rays.fetch_add(rays);
  };
  std::vector<std::thread> th;
  for (unsigned i = 1; i < nt; i++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    stats->paths = (uint64_t)W * H * (uint64_t)(s_end - s_begin);
    stats->rays = total_rays.load();
    stats->seconds = std::chrono::duration<double>(t1 - t0).count();
  }
  return 0;
}

void orc_resolve_u32(const float *image, int64_t n, uint32_t *out) {  // renderer.rs:112-120, color.rs:87-93
  auto chan = [](float c) -> uint32_t {
    float r = std::sqrt(c);
    // f32::clamp(0,1) keeps NaN; `as u32` maps NaN to 0 and saturates.
    if (r < 0.0f) r = 0.0f;
    if (r > 1.0f) r = 1.0f;
    float v = r * 255.0f;
    if (!(v > 0.0f)) return 0u;
    return (uint32_t)v;
  };
  for (int64_t i = 0; i < n; i++) {
    uint32_t r = chan(image[i * 3]), g = chan(image[i * 3 + 1]), b = chan(image[i * 3 + 2]);
    out[i] = (r << 16) | (g << 8) | b;
  }
}

void orc_camera_get_ray(const orc_camera *c, float u, float v, float origin[3], float dir[3]) {
  Ray r = camera_get_ray(*c, u, v);
  origin[0] = r.origin.x; origin[1] = r.origin.y; origin[2] = r.origin.z;
  dir[0] = r.direction.x; dir[1] = r.direction.y; dir[2] = r.direction.z;
}

int orc_scatter(const orc_material *m, const float ray_dir[3], const float position[3], const float normal[3],
                int front_face, const float u[4], float out_origin[3], float out_dir[3], float attenuation[3],
                float emitted[3]) {
  Ray ray_in = Ray::raw(Vec3(0, 0, 0), Vec3(ray_dir[0], ray_dir[1], ray_dir[2]));
  HitRecord hit;
  hit.position = Vec3(position[0], position[1], position[2]);
  hit.normal = Vec3(normal[0], normal[1], normal[2]);
  hit.front_face = front_face != 0;
  ArrayRng rng(u);
  Ray sc;
  Color att;
  Color e = mat_emitted(*m);
  emitted[0] = e.r; emitted[1] = e.g; emitted[2] = e.b;
  bool ok = mat_scatter(*m, ray_in, hit, rng, sc, att);
  if (ok) {
    out_origin[0] = sc.origin.x; out_origin[1] = sc.origin.y; out_origin[2] = sc.origin.z;
    out_dir[0] = sc.direction.x; out_dir[1] = sc.direction.y; out_dir[2] = sc.direction.z;
    attenuation[0] = att.r; attenuation[1] = att.g; attenuation[2] = att.b;
  }
  return ok ? 1 : 0;
}

int orc_mesh_bvh_info(const orc_scene *s, int object, int64_t *nodes, int64_t *leaves, int32_t *depth, uint8_t *dead,
                      int32_t *order) {
  if (object < 0 || (size_t)object >= s->objects.size()) return -1;
  const Mesh *m = dynamic_cast<const Mesh *>(s->objects[(size_t)object].get());
  if (!m) return -2;
  int64_t nn = 0, nl = 0;
  int md = 0;
  bvh_walk(*m->bvh, 0, false, nn, nl, md, dead);
  if (nodes) *nodes = nn;
  if (leaves) *leaves = nl;
  if (depth) *depth = md;
  if (order)
    for (size_t pos = 0; pos < m->order.size(); pos++) order[m->order[pos]] = (int32_t)pos;
  return 0;
}

void orc_chacha_stream(uint64_t seed, uint32_t *out, int n) {
  ChaChaRng r(seed);
  for (int i = 0; i < n; i++) out[i] = r.next_u32();
}

void orc_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]) {
  chacha_block(key, counter, stream, rounds, out);
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

int orc_div_panics(void) { return g_div_panics.load(); }

}  // extern "C"
