/* oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement of the hot path of
 * jackra1n/raytracer-rust (`render_scene` / `trace_ray`, src/renderer.rs:19-123, and everything
 * they call).  It exists to check the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs may load it.  The product library
 * (libptcore.so) never links, loads or calls it.
 *
 * Parity status: the reference ships no tests, no golden vectors and cannot be compiled in this
 * image (no cargo/rustc, crates not vendored).  The oracle is pinned only by the known answers
 * derived from the reference source (KA1..KA6, tests/test_oracle_known_answers.py).  Arithmetic
 * that lives in third-party crates (glam 0.30.3 Mat4 inverse / euler, rand 0.9.1 + rand_chacha
 * 0.9.0 StdRng stream, Rust's sort_unstable_by tie order) is restated from their published
 * algorithms and is PARITY UNPINNED.
 *
 * All structs are plain C, layouts deliberately identical to include/ptcore.h so one scene
 * description can be fed to both sides by the tests.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORC_MAT_LAMBERT = 0,         /* material.rs:29-71  AlbedoKind::Solid   */
  ORC_MAT_LAMBERT_CHECKER = 1, /* material.rs:29-71  AlbedoKind::Checked */
  ORC_MAT_METAL = 2,           /* material.rs:73-110 */
  ORC_MAT_DIELECTRIC = 3,      /* material.rs:112-167 */
  ORC_MAT_EMISSIVE = 4,        /* material.rs:169-192 */
  ORC_MAT_PLASTIC = 5,         /* tungsten/materials.rs:12-70 */
  ORC_MAT_ROUGH_CONDUCTOR = 6, /* tungsten/materials.rs:154-377 */
  ORC_MAT_NULL = 7             /* material.rs:229-252 */
};
enum { ORC_DIST_GGX = 0, ORC_DIST_BECKMANN = 1 };
enum { ORC_RNG_CHACHA = 0, /* per-row StdRng(seed=y), rejection sampling: the reference's stream */
       ORC_RNG_PHILOX = 1  /* counter-based, keyed (pixel,sample,bounce): the GPU's stream        */ };

typedef struct orc_material {
  int32_t type;
  float albedo[3];    /* lambert/metal/plastic/rough-conductor albedo, emissive colour, checker on_color */
  float off_color[3]; /* checker off_color */
  float inv_scale;    /* checker inv_scale (CheckerTexture::new already applied) */
  float fuzz;         /* metal fuzz, already clamped to [0,1] (Metal::new) */
  float ior;          /* dielectric refractive_index / plastic ior */
  float roughness;    /* rough conductor, already max(0.01) (RoughConductor::new) */
  float eta[3];       /* MetalType::ior_k().0 */
  float k[3];         /* MetalType::ior_k().1 */
  int32_t distribution;
} orc_material;

typedef struct orc_camera { /* camera.rs:4-11 */
  float position[3], forward[3], right[3], true_up[3];
  float half_width, half_height;
} orc_camera;

typedef struct orc_settings {
  int32_t width, height;
  int32_t spp;          /* samples per pixel (total) */
  int32_t max_depth;    /* renderer.rs:101: total segments including the camera ray */
  uint64_t seed;        /* Philox key; ignored in ChaCha mode (seed = row index, renderer.rs:91) */
  int32_t rng_mode;     /* ORC_RNG_* */
  int32_t sample_begin; /* Philox mode only: render samples [sample_begin, sample_end) */
  int32_t sample_end;   /* 0,0 = all */
  int32_t threads;      /* 0 = hardware concurrency */
} orc_settings;

typedef struct orc_hit {
  int32_t object;    /* index in object_list, -1 = miss */
  int32_t triangle;  /* index in Mesh.triangles (after the degenerate filter), -1 for analytic prims */
  float t;
  float position[3];
  float normal[3];
  int32_t front_face;
  int32_t material;
} orc_hit;

typedef struct orc_stats {
  uint64_t paths;
  uint64_t rays; /* trace_ray calls with depth > 0 */
  double seconds; /* render loop only (renderer.rs:82,109) */
} orc_stats;

typedef struct orc_scene orc_scene;

orc_scene *orc_scene_create(void);
void orc_scene_destroy(orc_scene *);
int orc_scene_add_material(orc_scene *, const orc_material *);
int orc_scene_add_sphere(orc_scene *, const float center[3], float radius, int material);
int orc_scene_add_plane(orc_scene *, const float p1[3], const float normal[3], int material);
int orc_scene_add_quad(orc_scene *, const float base[3], const float e0[3], const float e1[3],
                       const float normal[3], float d, float inv_e0_len_sq, float inv_e1_len_sq, int material);
int orc_scene_add_cube(orc_scene *, const float o2w[16], const float w2o[16], int material);
/* tris: n x 12 floats = v0,v1,v2,normal (object space, already degenerate-filtered) */
int orc_scene_add_mesh(orc_scene *, const float *tris, int64_t n, const float o2w[16], const float w2o[16], int material);
int orc_scene_set_sky_hdr(orc_scene *, const float *rgb, int w, int h);

/* HittableList::hit (hittable.rs:46-57) on caller-provided rays.  Directions are used as given
 * (callers pass what Ray::new would hold, i.e. already normalised). */
int orc_intersect(const orc_scene *, const float *origins, const float *dirs, int64_t n, float t_min, float t_max,
                  orc_hit *out);
/* image: W*H*3 floats, linear mean radiance (= image_data, renderer.rs:83,103) */
int orc_render(const orc_scene *, const orc_camera *, const orc_settings *, float *image, orc_stats *stats);
/* renderer.rs:112-120 + color.rs:87-93 */
void orc_resolve_u32(const float *image, int64_t n_pixels, uint32_t *out);
/* camera.rs:33-42 */
void orc_camera_get_ray(const orc_camera *, float u, float v, float origin[3], float dir[3]);

/* Material::scatter with explicit uniforms (u[0..3] are consumed in the Philox-mode order documented in
 * DESIGN.md).  Returns 1 if scattered, 0 if absorbed.  Also returns emitted(). */
int orc_scatter(const orc_material *, const float ray_dir[3], const float position[3], const float normal[3],
                int front_face, const float u[4], float out_origin[3], float out_dir[3], float attenuation[3],
                float emitted[3]);

/* Reference BVH facts for mesh object `object` (bvh.rs:15-76): node / leaf counts, depth (root = 0),
 * per-triangle flags: dead[i]=1 if some ancestor node (leaf included) has zero extent on an axis
 * (aabb.rs:40 makes such nodes unhittable), order[i] = position of triangle i in DFS leaf order. */
int orc_mesh_bvh_info(const orc_scene *, int object, int64_t *nodes, int64_t *leaves, int32_t *depth,
                      uint8_t *dead, int32_t *order);

/* RNG known-answer hooks */
void orc_chacha_stream(uint64_t seed, uint32_t *out, int n);          /* StdRng::seed_from_u64(seed).next_u32() x n */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* raw ChaCha block (rounds = 12 for StdRng, 20 for the RFC 7539 known answer) */
void orc_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]);
/* number of times the reference would have panicked in `Vec3 / f32` (vec3.rs:120-122) so far */
int orc_div_panics(void);

#ifdef __cplusplus
}
#endif
#endif
