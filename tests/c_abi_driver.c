/* c_abi_driver.c — a plain-C caller of libptcore.so: what the reference's Rust host does through `extern "C"` after the
 * patch of INTEGRATION.md section 4 (the body of render_scene, /root/reference/src/renderer.rs:67-123), written in the one
 * other language this image can compile.  No Python, no torch, no C++: gcc + include/ptcore.h only.
 *
 *   c_abi_driver <scene.bin> <out.u32> [device]
 *
 * scene.bin is the flat dump of a scene description in object_list order that tests/test_c_abi_driver.py writes (the
 * values `Hittable::describe()` / `Material::describe()` would forward: already-derived matrices, quad normal / d /
 * inverse edge lengths, triangle normals).  The program replays it through ptc_scene_add_* in list order, commits, calls
 * ptc_render_u32 and writes render_scene's Vec<u32> to out.u32.  Exit code: 0 ok, 2 = PTC_E_CUDA at commit (no device:
 * the library has no CPU fallback), 1 = anything else.
 *
 * File layout (little endian): u32 magic 'PTC1'; ptc_camera; i32 width, height, spp, max_depth; u64 seed;
 *   i32 n_materials; ptc_material[n]; i32 n_objects; then per object: i32 type, i32 material, and
 *   sphere: f32 center[3], radius | plane: f32 p1[3], n[3] | quad: f32 base[3], e0[3], e1[3], n[3], d, inv0, inv1 |
 *   cube: f32 o2w[16], w2o[16] | mesh: f32 o2w[16], w2o[16], i64 n_tris, f32 tris[n_tris][12];
 *   i32 sky_w, sky_h; f32 sky[w*h*3] (w = 0: none). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptcore.h"

static FILE *in;
static void rd(void *p, size_t n) {
  if (fread(p, 1, n, in) != n) {
    fprintf(stderr, "c_abi_driver: short read\n");
    exit(1);
  }
}
#define CHECK(call)                                                      \
  do {                                                                   \
    int rc_ = (call);                                                    \
    if (rc_ < 0) {                                                       \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ptc_last_error());   \
      return rc_ == PTC_E_CUDA ? 2 : 1;                                  \
    }                                                                    \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 3) return 1;
  in = fopen(argv[1], "rb");
  if (!in) return 1;
  const int device = argc > 3 ? atoi(argv[3]) : 0;
  uint32_t magic;
  rd(&magic, 4);
  if (magic != 0x31435450u) return 1;
  if (ptc_abi_version() != PTC_ABI_VERSION) return 1;
  ptc_camera cam;
  rd(&cam, sizeof cam);
  ptc_render_settings st;
  memset(&st, 0, sizeof st);
  rd(&st.width, 4), rd(&st.height, 4), rd(&st.spp, 4), rd(&st.max_depth, 4), rd(&st.seed, 8);
  ptc_scene *scene = ptc_scene_create();
  if (!scene) return 1;
  int32_t n;
  rd(&n, 4);
  for (int32_t i = 0; i < n; i++) {
    ptc_material m;
    rd(&m, sizeof m);
    CHECK(ptc_scene_add_material(scene, &m));
  }
  rd(&n, 4);
  for (int32_t i = 0; i < n; i++) { /* the ORDER of these calls is Scene.object_list order (hittable.rs:50-55) */
    int32_t type, material;
    float f[32];
    rd(&type, 4), rd(&material, 4);
    switch (type) {
      case 0: rd(f, 16); CHECK(ptc_scene_add_sphere(scene, f, f[3], material)); break;
      case 1: rd(f, 24); CHECK(ptc_scene_add_plane(scene, f, f + 3, material)); break;
      case 2: rd(f, 60); CHECK(ptc_scene_add_quad(scene, f, f + 3, f + 6, f + 9, f[12], f[13], f[14], material)); break;
      case 3: rd(f, 128); CHECK(ptc_scene_add_cube(scene, f, f + 16, material)); break;
      case 4: {
        int64_t nt;
        rd(f, 128), rd(&nt, 8);
        float *tris = (float *)malloc((size_t)nt * 48);
        if (!tris) return 1;
        rd(tris, (size_t)nt * 48);
        CHECK(ptc_scene_add_mesh(scene, tris, nt, f, f + 16, material));
        free(tris); /* inputs are copied during the add_* call */
        break;
      }
      default: return 1;
    }
  }
  int32_t sw, sh;
  rd(&sw, 4), rd(&sh, 4);
  if (sw > 0) {
    float *sky = (float *)malloc((size_t)sw * sh * 12);
    if (!sky) return 1;
    rd(sky, (size_t)sw * sh * 12);
    CHECK(ptc_scene_set_sky_hdr(scene, sky, sw, sh));
    free(sky);
  }
  fclose(in);
  CHECK(ptc_scene_commit(scene, device));
  uint32_t *out = (uint32_t *)malloc((size_t)st.width * st.height * 4);
  ptc_stats stats;
  CHECK(ptc_render_u32(scene, &cam, &st, out, &stats)); /* render_scene(&scene, &camera, &settings) -> Vec<u32> */
  FILE *o = fopen(argv[2], "wb");
  if (!o || fwrite(out, 4, (size_t)st.width * st.height, o) != (size_t)st.width * st.height) return 1;
  fclose(o);
  printf("{\"paths\": %llu, \"rays\": %llu, \"render_ms\": %.3f, \"launches\": %llu}\n", (unsigned long long)stats.paths,
         (unsigned long long)stats.rays, stats.render_ms, (unsigned long long)stats.kernel_launches);
  free(out);
  ptc_scene_destroy(scene);
  return 0;
}
