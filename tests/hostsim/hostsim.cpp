// hostsim.cpp — TEST-ONLY host compilation of the device headers (raytracer-rust_b200/csrc/pt_*.h).
//
// The build container has no GPU.  The intersection / traversal / BSDF / RNG code of the product is written once
// as __host__ __device__ inline functions; this file compiles those SAME headers with g++ (-ffp-contract=off) so
// `pytest -m "not gpu"` can compare them with the oracle before a GPU is involved.  It is not part of the product:
// libptcore.so never links or loads it, it exports `sim_*` symbols only, and nothing outside tests/ may use it.
// It contains no renderer — rendering exists only as CUDA kernels (csrc/ptcore.cu).
#include <cstdint>
#include <cstring>
#include <string>

#include "../../raytracer-rust_b200/csrc/pt_bsdf.h"
#include "../../raytracer-rust_b200/csrc/pt_philox.h"
#include "../../raytracer-rust_b200/csrc/pt_prims.h"
#include "../../raytracer-rust_b200/csrc/pt_scene_host.h"

using namespace pt;

struct sim_scene {
  HostScene hs;
  std::vector<DMesh> dmeshes;
  DScene ds;
  bool committed = false;
};

static thread_local std::string g_err;

extern "C" {

const char *sim_last_error(void) { return g_err.c_str(); }
sim_scene *sim_scene_create(void) { return new sim_scene(); }
void sim_scene_destroy(sim_scene *s) { delete s; }

#define SIM_TRY(expr)          \
  try {                        \
    return (expr);             \
  } catch (std::exception & e) { \
    g_err = e.what();          \
    return -1;                 \
  }

int sim_scene_add_material(sim_scene *s, const ptc_material *m) { SIM_TRY(s->hs.add_material(m)) }
int sim_scene_add_sphere(sim_scene *s, const float c[3], float r, int mat) { SIM_TRY(s->hs.add_sphere(c, r, mat)) }
int sim_scene_add_plane(sim_scene *s, const float p[3], const float n[3], int mat) { SIM_TRY(s->hs.add_plane(p, n, mat)) }
int sim_scene_add_quad(sim_scene *s, const float b[3], const float e0[3], const float e1[3], const float n[3], float d,
                       float i0, float i1, int mat) {
  SIM_TRY(s->hs.add_quad(b, e0, e1, n, d, i0, i1, mat))
}
int sim_scene_add_cube(sim_scene *s, const float o2w[16], const float w2o[16], int mat) {
  SIM_TRY(s->hs.add_cube(o2w, w2o, mat))
}
int sim_scene_add_mesh(sim_scene *s, const float *tris, int64_t n, const float o2w[16], const float w2o[16], int mat) {
  SIM_TRY(s->hs.add_mesh(tris, n, o2w, w2o, mat))
}
int sim_scene_set_sky_hdr(sim_scene *s, const float *rgb, int32_t w, int32_t h) {
  try {
    s->hs.set_sky(rgb, w, h);
    return 0;
  } catch (std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

int sim_scene_commit(sim_scene *s, int /*device*/) {
  try {
    s->hs.build_all();
    s->dmeshes.clear();
    for (auto &m : s->hs.meshes) {
      const DMesh d = make_dmesh(*m, reinterpret_cast<const float4 *>(m->nodes.data()),
                                 reinterpret_cast<const float4 *>(m->tri48.data()), m->normals.data());
      s->dmeshes.push_back(d);
    }
    s->ds.objects = s->hs.objects.data();
    s->ds.materials = s->hs.materials.data();
    s->ds.meshes = s->dmeshes.data();
    s->ds.sky = s->hs.sky.empty() ? nullptr : s->hs.sky.data();
    s->ds.lights = s->hs.lights.data();
    s->ds.n_lights = (int32_t)s->hs.lights.size();
    s->ds.n_objects = (int32_t)s->hs.objects.size();
    s->ds.n_materials = (int32_t)s->hs.materials.size();
    s->ds.n_meshes = (int32_t)s->dmeshes.size();
    s->ds.sky_w = s->hs.sky_w;
    s->ds.sky_h = s->hs.sky_h;
    s->committed = true;
    return 0;
  } catch (std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// for tests/hostsim/wfsim.cpp (the SIMT cost model)
const DScene *sim_scene_dscene(const sim_scene *s) { return s->committed ? &s->ds : nullptr; }

int sim_scene_mesh_info(const sim_scene *s, int object, ptc_mesh_info *info, uint8_t *dead, int32_t *order) {
  if (object < 0 || (size_t)object >= s->hs.objects.size() || s->hs.objects[object].type != OBJ_MESH) return -1;
  const MeshBuild &m = *s->hs.meshes[s->hs.objects[object].mesh];
  if (!m.built) return -4;
  if (info) {
    info->triangles = m.n;
    info->live_triangles = m.live;
    info->ref_nodes = m.ref_nodes;
    info->ref_leaves = m.ref_leaves;
    info->ref_depth = m.ref_depth;
    info->wide_nodes = (int64_t)m.nodes.size();
    info->wide_depth = m.wide_depth;
    info->node_bytes = (int64_t)m.nodes.size() * 80;
    info->triangle_bytes = (int64_t)m.tri48.size() * 48;
  }
  if (dead) memcpy(dead, m.dead.data(), (size_t)m.n);
  if (order) memcpy(order, m.order.data(), (size_t)m.n * 4);
  return 0;
}

int sim_intersect(sim_scene *s, const float *origins, const float *dirs, int64_t n, float t_min, float t_max,
                  ptc_hit *out, ptc_stats *stats) {
  if (!s->committed) return -4;
  TraversalCounters ctr{0, 0, 0};
  uint64_t nodes = 0, tris = 0, mesh_rays = 0;
  for (int64_t i = 0; i < n; i++) {
    Ray r{v3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]), v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2])};
    Hit h;
    memset(&h, 0, sizeof(h));
    ctr = TraversalCounters{0, 0, 0};
    bool hit = scene_hit<true>(s->ds, r, t_min, t_max, h, &ctr);
    nodes += ctr.nodes;
    tris += ctr.tris;
    mesh_rays += ctr.mesh_rays;
    ptc_hit &o = out[i];
    memset(&o, 0, sizeof(o));
    if (hit) {
      o.object = h.object;
      o.triangle = h.triangle;
      o.t = h.t;
      o.position[0] = h.px, o.position[1] = h.py, o.position[2] = h.pz;
      o.normal[0] = h.nx, o.normal[1] = h.ny, o.normal[2] = h.nz;
      o.front_face = h.front_face;
      o.material = h.material;
    } else {
      o.object = -1;
      o.triangle = -1;
      o.material = -1;
    }
  }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->rays = (uint64_t)n;
    stats->nodes_visited = nodes;
    stats->tris_tested = tris;
    stats->mesh_rays = mesh_rays;
  }
  return 0;
}

int sim_primary_rays(sim_scene *, const ptc_camera *cam, const ptc_render_settings *st, int32_t sample, float *out_o,
                     float *out_d) {
  DCamera c;
  memcpy(&c, cam, sizeof(c));
  for (int y = 0; y < st->height; y++)
    for (int x = 0; x < st->width; x++) {
      const uint32_t pixel = (uint32_t)(y * st->width + x);
      Uniforms4 j = philox_uniforms(st->seed, pixel, (uint32_t)sample, 0xffffffffu, 0u);
      const float u = ((float)x + j.u[0]) / (float)st->width;
      const float v = ((float)y + j.u[1]) / (float)st->height;
      Ray r = camera_get_ray(c, u, v);
      out_o[pixel * 3 + 0] = r.o.x, out_o[pixel * 3 + 1] = r.o.y, out_o[pixel * 3 + 2] = r.o.z;
      out_d[pixel * 3 + 0] = r.d.x, out_d[pixel * 3 + 1] = r.d.y, out_d[pixel * 3 + 2] = r.d.z;
    }
  return 0;
}

int sim_scatter(sim_scene *s, int material, const float *ray_dirs, const float *positions, const float *normals,
                const int32_t *front_face, const float *u4, int64_t n, int32_t *scattered, float *out_origin,
                float *out_dir, float *attenuation, float *emitted) {
  if (material < 0 || (size_t)material >= s->hs.materials.size()) return -1;
  const DMaterial &m = s->hs.materials[material];
  for (int64_t i = 0; i < n; i++) {
    V3 d = v3(ray_dirs[i * 3], ray_dirs[i * 3 + 1], ray_dirs[i * 3 + 2]);
    V3 p = v3(positions[i * 3], positions[i * 3 + 1], positions[i * 3 + 2]);
    V3 nn = v3(normals[i * 3], normals[i * 3 + 1], normals[i * 3 + 2]);
    Ray sc{v3(0, 0, 0), v3(0, 0, 0)};
    V3 att = v3(0, 0, 0);
    V3 e = mat_emitted(m);
    bool ok = mat_scatter(m, d, p, nn, front_face[i] != 0, u4 + i * 4, sc, att);
    scattered[i] = ok ? 1 : 0;
    out_origin[i * 3] = sc.o.x, out_origin[i * 3 + 1] = sc.o.y, out_origin[i * 3 + 2] = sc.o.z;
    out_dir[i * 3] = sc.d.x, out_dir[i * 3 + 1] = sc.d.y, out_dir[i * 3 + 2] = sc.d.z;
    attenuation[i * 3] = att.x, attenuation[i * 3 + 1] = att.y, attenuation[i * 3 + 2] = att.z;
    emitted[i * 3] = e.x, emitted[i * 3 + 1] = e.y, emitted[i * 3 + 2] = e.z;
  }
  return 0;
}

int sim_philox(sim_scene *, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  U4 r = philox4x32_10(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
  out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
  return 0;
}

int sim_resolve_u32(sim_scene *, const float *rgb, int64_t n, float scale, uint32_t *out) {
  for (int64_t i = 0; i < n; i++) out[i] = resolve_pixel(rgb[i * 3] * scale, rgb[i * 3 + 1] * scale, rgb[i * 3 + 2] * scale);
  return 0;
}

}  // extern "C"
