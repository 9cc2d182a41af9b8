// wfsim.cpp — TEST/TOOLING ONLY: a SIMT cost model of the wavefront kernels, run on the CPU.
//
// A gpurun round trip costs minutes; whether a change of *scheduling* (which lanes of a warp execute which step, how
// many exact primitive tests a warp runs for how many useful lanes) pays off can be answered here in seconds.  This
// file replays ONE segment of the segmented wavefront (csrc/pt_wavefront.cuh: same chunking into 256-ray blocks and
// 32-lane warps, same class-sorted survivor order out of the shade stage, same regeneration order) with the product's
// own device functions compiled by g++, and counts warp-level instruction slots under alternative policies:
//
//   extend pre/post : every lane runs every exact primitive test (v1)  vs  per-lane conservative world-box cull  vs
//                     block-wide compaction of the lanes that pass the cull (object rounds)
//   traverse        : the persistent-warp loop of stage_traverse under different step-selection / refill policies
//
// It is a model (fixed per-step instruction weights), used to rank designs before they are measured on the GPU; the
// numbers that count are the CUDA-event and ncu measurements under profiles/.  Exports sim_* symbols only.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../raytracer-rust_b200/csrc/pt_bsdf.h"
#include "../../raytracer-rust_b200/csrc/pt_philox.h"
#include "../../raytracer-rust_b200/csrc/pt_prims.h"
#include "../../raytracer-rust_b200/csrc/pt_scene_host.h"

using namespace pt;

struct sim_scene;  // hostsim.cpp
extern "C" const DScene *sim_scene_dscene(const sim_scene *s);

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kHitBit = 0x80000000u, kFrontBit = 0x40000000u, kMatMask = 0x3fffffffu;

struct PathState {
  V3 o, d, beta;
  uint32_t pixel, sample, bounce;
  // hit record
  V3 p, n;
  float t;
  uint32_t bits;
};

struct Task {
  uint32_t ray;
  int object;
  V3 o, d;
  float closest;
  float res_t;
  uint32_t res_tri;
};

struct Box {
  float lo[3], hi[3];
};

// conservative world-space box of an object (what a commit-time precomputation would store)
Box world_box(const DScene &sc, const DObject &ob) {
  Box b;
  for (int a = 0; a < 3; a++) b.lo[a] = INFINITY, b.hi[a] = -INFINITY;
  auto grow = [&](V3 p) {
    const float v[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; a++) b.lo[a] = std::min(b.lo[a], v[a]), b.hi[a] = std::max(b.hi[a], v[a]);
  };
  if (ob.type == OBJ_SPHERE) {
    const float r = fabsf(ob.f[3]);
    grow(v3(ob.f[0] - r, ob.f[1] - r, ob.f[2] - r));
    grow(v3(ob.f[0] + r, ob.f[1] + r, ob.f[2] + r));
  } else if (ob.type == OBJ_QUAD) {
    const V3 base = v3(ob.f[0], ob.f[1], ob.f[2]), e0 = v3(ob.f[3], ob.f[4], ob.f[5]), e1 = v3(ob.f[6], ob.f[7], ob.f[8]);
    grow(base), grow(base + e0), grow(base + e1), grow(base + e0 + e1);
  } else if (ob.type == OBJ_CUBE || ob.type == OBJ_MESH) {
    float lo[3] = {-0.5f, -0.5f, -0.5f}, hi[3] = {0.5f, 0.5f, 0.5f};
    if (ob.type == OBJ_MESH)
      for (int a = 0; a < 3; a++) lo[a] = sc.meshes[ob.mesh].root_lo[a], hi[a] = sc.meshes[ob.mesh].root_hi[a];
    for (int c = 0; c < 8; c++) grow(mat_point(ob.f + 16, v3(c & 1 ? hi[0] : lo[0], c & 2 ? hi[1] : lo[1], c & 4 ? hi[2] : lo[2])));
  } else {
    for (int a = 0; a < 3; a++) b.lo[a] = -INFINITY, b.hi[a] = INFINITY;
  }
  for (int a = 0; a < 3; a++) {
    const float pad = 1e-3f * std::max(1.0f, std::max(fabsf(b.lo[a]), fabsf(b.hi[a])));
    b.lo[a] -= pad, b.hi[a] += pad;
  }
  return b;
}
bool box_may_hit(const Box &b, const Ray &r, float t_min, float t_max) {
  const float ix = 1.0f / r.d.x, iy = 1.0f / r.d.y, iz = 1.0f / r.d.z;
  const float ax = (b.lo[0] - r.o.x) * ix, bx = (b.hi[0] - r.o.x) * ix;
  const float ay = (b.lo[1] - r.o.y) * iy, by = (b.hi[1] - r.o.y) * iy;
  const float az = (b.lo[2] - r.o.z) * iz, bz = (b.hi[2] - r.o.z) * iz;
  const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t_min));
  const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t_max * 1.0001f));
  return tn <= tf;
}

// instruction weights (warp instructions per executed step; from the SASS of the v1 kernels, rounded)
struct Weights {
  double cull = 14, cull_setup = 12;
  double reject[5] = {30, 20, 40, 110, 190};   // sphere, plane, quad, cube, mesh (object ray + root test)
  double hit_extra[5] = {60, 15, 20, 170, 0};  // additional when a lane accepts the hit
  double node = 360, tri = 120, turn = 60, refill = 70, finish = 12;  // ncu, C2 (profiles/r1_v3_*): per executed step
  double barrier = 12;  // per warp per block barrier (issue + expected skew), a guess
};

struct ExtendModel {
  double now = 0, lanecull = 0, blockcompact = 0, masksort = 0, pairs = 0, n_pairs = 0;  // warp-instruction slots
  double tests = 0, tests_pass_cull = 0, hits = 0;
  double warp_obj = 0, warp_obj_any_pass = 0;
};

struct ScanRec {  // what one lane did for one object of its scan
  uint8_t reached, pass, hit, pass_inf;
};

struct TravModel {
  double cost = 0, ideal = 0;
  double turns = 0, node_exec = 0, tri_exec = 0, node_lanes = 0, tri_lanes = 0, refills = 0, hits = 0;
  double hist_nodes[16] = {0}, hist_tris[16] = {0};
};

struct TravPolicy {
  int refill_lanes = 8;
  int mode = 0;      // 0 = v1: node step and tri step every turn for whoever is ready
                     // 1 = vote: run the step with more ready lanes; the other one only if it has >= thr lanes ready
                     // 2 = tri-first drain: tris whenever >= thr_tri lanes have one, nodes otherwise
  int thr_node = 0, thr_tri = 0;
  int tri_groups = 2;  // pending leaf groups a lane can hold before its node steps have to wait (v1: 2)
};

struct LaneState {
  bool active = false;
  TravState s;
  DMesh mesh;
  uint2 stack[kTraversalStack];
  int sp = 0;
  uint32_t task = 0;
  uint32_t nsteps = 0, tsteps = 0;
  // extra leaf groups beyond tg/tg2 for tri_groups > 2 (model only)
  std::vector<uint2> extra;
};

// one block (8 warps sharing the task cursor) under a policy; fills the task results, accumulates the model
void run_traverse(const DScene &sc, std::vector<Task> &tasks, float t_min, const TravPolicy &pol, const Weights &w, TravModel &tm) {
  const uint32_t n = (uint32_t)tasks.size();
  if (n == 0) return;
  uint32_t cursor = 0;
  const int n_warps = kBlock / 32;
  std::vector<LaneState> lanes((size_t)n_warps * 32);
  std::vector<char> exhausted(n_warps, 0), done(n_warps, 0);
  int remaining = n_warps;
  TraversalCounters tc{0, 0, 0};
  while (remaining > 0) {
    for (int wi = 0; wi < n_warps; wi++) {
      if (done[wi]) continue;
      LaneState *L = &lanes[(size_t)wi * 32];
      tm.turns += 1;
      tm.cost += w.turn;
      int idle = 0;
      for (int l = 0; l < 32; l++) idle += !L[l].active;
      if (!exhausted[wi] && (idle == 32 || idle >= pol.refill_lanes)) {
        tm.refills += 1;
        tm.cost += w.refill;
        const uint32_t base = cursor;
        cursor += (uint32_t)idle;
        uint32_t j = base;
        for (int l = 0; l < 32; l++) {
          if (L[l].active) continue;
          if (j < n) {
            Task &t = tasks[j];
            const DMesh &gm = sc.meshes[sc.objects[t.object].mesh];
            L[l].mesh = gm;
            trav_begin(L[l].s, t.o, t.d, t_min, t.closest);
            L[l].sp = 0;
            L[l].task = j;
            L[l].active = true;
            L[l].extra.clear();
            L[l].nsteps = 0, L[l].tsteps = 0;
          }
          j++;
        }
        if (base + (uint32_t)idle >= n) exhausted[wi] = 1;
      }
      int n_active = 0;
      for (int l = 0; l < 32; l++) n_active += L[l].active;
      if (n_active == 0) {
        if (exhausted[wi]) {
          done[wi] = 1;
          remaining--;
        }
        continue;
      }
      // pop / finish
      for (int l = 0; l < 32; l++) {
        LaneState &a = L[l];
        if (!a.active) continue;
        if (!trav_has_node(a.s)) {
          if (a.sp > 0) a.s.ng = a.stack[--a.sp];
          else if (!trav_has_tri(a.s)) {
            tasks[a.task].res_t = a.s.best_t;
            tasks[a.task].res_tri = a.s.best_tri;
            a.active = false;
            tm.cost += w.finish / 32.0;
            tm.hist_nodes[std::min<uint32_t>(a.nsteps, 15)] += 1;
            tm.hist_tris[std::min<uint32_t>(a.tsteps, 15)] += 1;
            if (a.s.best_tri != 0xffffffffu) tm.hits += 1;
          }
        }
      }
      int node_ready = 0, tri_ready = 0;
      for (int l = 0; l < 32; l++) {
        LaneState &a = L[l];
        if (!a.active) continue;
        const bool room = pol.tri_groups <= 2 ? a.s.tg2.y == 0u : (int)a.extra.size() + 2 < pol.tri_groups || a.s.tg2.y == 0u;
        if (trav_has_node(a.s) && room) node_ready++;
        if (trav_has_tri(a.s)) tri_ready++;
      }
      bool do_node = node_ready > 0, do_tri = tri_ready > 0;
      if (pol.mode == 1) {
        if (node_ready >= tri_ready) do_tri = do_tri && (tri_ready >= pol.thr_tri || node_ready == 0);
        else do_node = do_node && (node_ready >= pol.thr_node || tri_ready == 0);
      } else if (pol.mode == 2) {
        if (tri_ready >= pol.thr_tri || node_ready == 0) do_node = do_node && node_ready >= pol.thr_node;
        else do_tri = false;
        if (!do_node && !do_tri) do_node = node_ready > 0, do_tri = !do_node;
      }
      if (do_node) {
        tm.node_exec += 1;
        tm.cost += w.node;
        for (int l = 0; l < 32; l++) {
          LaneState &a = L[l];
          if (!a.active || !trav_has_node(a.s)) continue;
          if (a.s.tg2.y != 0u) {
            if ((int)a.extra.size() + 2 >= pol.tri_groups) continue;
            // model of a deeper pending-group queue: park tg2 in `extra`
            a.extra.push_back(a.s.tg2);
            a.s.tg2 = make_uint2(0u, 0u);
          }
          trav_node<true>(a.mesh, a.s, a.stack, a.sp, &tc);
          tm.node_lanes += 1;
          a.nsteps++;
          tm.ideal += w.node / 32.0;
        }
      }
      if (do_tri) {
        tm.tri_exec += 1;
        tm.cost += w.tri;
        for (int l = 0; l < 32; l++) {
          LaneState &a = L[l];
          if (!a.active || !trav_has_tri(a.s)) continue;
          trav_tri<true>(a.mesh, a.s, &tc);
          if (a.s.tg2.y == 0u && !a.extra.empty()) {  // refill tg2 (or tg) from the modelled queue, oldest first
            if (a.s.tg.y == 0u) a.s.tg = a.extra.front();
            else a.s.tg2 = a.extra.front();
            a.extra.erase(a.extra.begin());
          }
          tm.tri_lanes += 1;
          a.tsteps++;
          tm.ideal += w.tri / 32.0;
        }
      }
    }
  }
}

}  // namespace

extern "C" {

struct sim_model_params {
  int32_t width, height, max_depth, iterations, cap, n_policies;
  uint64_t seed;
  int32_t policy[8][5];  // refill_lanes, mode, thr_node, thr_tri, tri_groups
};

// out: doubles, see tools/simt_model.py for the layout
int sim_wavefront_model(sim_scene *ss, const ptc_camera *cam, const sim_model_params *mp, double *out, int n_out) {
  const DScene &sc = *sim_scene_dscene(ss);
  DCamera dcam;
  memcpy(&dcam, cam, sizeof(dcam));
  const Weights w;
  const float t_min = kEps;
  std::vector<Box> boxes;
  for (int k = 0; k < sc.n_objects; k++) boxes.push_back(world_box(sc, sc.objects[k]));

  const int tiles_x = (mp->width + 31) / 32, tiles_y = (mp->height + 31) / 32;
  const uint64_t per_sample = (uint64_t)tiles_x * tiles_y * 1024ull;
  const uint64_t rows = per_sample >> 5;
  auto gcd = [](uint64_t a, uint64_t b) {
    while (b) {
      const uint64_t t = a % b;
      a = b, b = t;
    }
    return a;
  };
  uint64_t row_mult = (0x9E3779B1ull % rows) | 1ull;
  while (gcd(row_mult, rows) != 1) row_mult += 2;

  std::vector<PathState> pool((size_t)mp->cap);
  uint32_t n = 0;
  uint64_t next_path = 0;
  ExtendModel em;
  TravModel tm[8];
  double rays = 0, n_tasks = 0, shade_lane_eff_num = 0, shade_lane_eff_den = 0;
  double sh_sum[2] = {0, 0}, sh_sync[2] = {0, 0};  // window 256 / 2048: sum of group costs, barrier-synchronous cost

  auto regenerate = [&]() {
    while (n < (uint32_t)mp->cap) {
      const uint64_t p = next_path++;
      const uint32_t s = (uint32_t)(p / per_sample);
      const uint32_t r0 = (uint32_t)(p - (uint64_t)s * per_sample);
      const uint32_t row = (uint32_t)(((uint64_t)(r0 >> 5) * row_mult) % rows);
      const uint32_t r = (row << 5) | (r0 & 31u);
      const uint32_t tile = r >> 10, in_tile = r & 1023u;
      const int x = (int)((tile % (uint32_t)tiles_x) * 32u + (in_tile & 31u));
      const int y = (int)((tile / (uint32_t)tiles_x) * 32u + (in_tile >> 5));
      if (x >= mp->width || y >= mp->height) continue;
      const uint32_t pixel = (uint32_t)(y * mp->width + x);
      const Uniforms4 jit = philox_uniforms(mp->seed, pixel, s, 0xffffffffu, 0u);
      const float u = ((float)x + jit.u[0]) / (float)mp->width;
      const float v = ((float)y + jit.u[1]) / (float)mp->height;
      const Ray ray = camera_get_ray(dcam, u, v);
      PathState &ps = pool[n++];
      ps.o = ray.o, ps.d = ray.d, ps.beta = v3(1, 1, 1);
      ps.pixel = pixel, ps.sample = s, ps.bounce = 0;
    }
  };
  regenerate();

  std::vector<std::vector<ScanRec>> recs;  // per ray of the chunk, per object
  for (int it = 0; it < mp->iterations; it++) {
    rays += n;
    std::vector<Task> tasks;
    std::vector<float> closest(n, INFINITY);
    // ---- pre: scan from object 0; post: scan from k + 1.  Both are "scan passes" for the extend model.
    struct Pending {
      uint32_t ray;
      int k_begin;
    };
    std::vector<Pending> todo(n);
    for (uint32_t i = 0; i < n; i++) todo[i] = Pending{i, 0}, pool[i].bits = 0u;
    bool first_pass = true;
    while (!todo.empty()) {
      std::vector<Task> parked;
      for (size_t c0 = 0; c0 < todo.size(); c0 += kBlock) {
        const size_t cn = std::min((size_t)kBlock, todo.size() - c0);
        recs.assign(cn, std::vector<ScanRec>((size_t)sc.n_objects, ScanRec{0, 0, 0, 0}));
        for (size_t j = 0; j < cn; j++) {
          const Pending pd = todo[c0 + j];
          PathState &ps = pool[pd.ray];
          const Ray ray{ps.o, ps.d};
          float &cl = closest[pd.ray];
          for (int k = pd.k_begin; k < sc.n_objects; k++) {
            const DObject *ob = sc.objects + k;
            ScanRec &rc = recs[j][(size_t)k];
            rc.reached = 1;
            rc.pass = box_may_hit(boxes[(size_t)k], ray, t_min, cl) ? 1 : 0;
            rc.pass_inf = box_may_hit(boxes[(size_t)k], ray, t_min, INFINITY) ? 1 : 0;
            Hit tmp;
            tmp.triangle = -1;
            bool hit = false;
            if (ob->type == OBJ_MESH) {
              const MeshRay mr = mesh_object_ray(ob->f, ray);
              if (!mesh_root_may_hit(sc.meshes[ob->mesh], mr, t_min, cl)) continue;
              rc.hit = 1;
              Task t;
              t.ray = pd.ray, t.object = k, t.o = mr.o, t.d = mr.d, t.closest = cl, t.res_t = 0, t.res_tri = 0xffffffffu;
              parked.push_back(t);
              break;
            }
            if (ob->type == OBJ_SPHERE) hit = hit_sphere(ob->f, ray, t_min, cl, tmp);
            else if (ob->type == OBJ_QUAD) hit = hit_quad(ob->f, ray, t_min, cl, tmp);
            else if (ob->type == OBJ_CUBE) hit = hit_cube(ob->f, ray, t_min, cl, tmp);
            else hit = hit_plane(ob->f, ray, t_min, cl, tmp);
            if (hit) {
              rc.hit = 1;
              cl = tmp.t;
              ps.t = tmp.t, ps.p = v3(tmp.px, tmp.py, tmp.pz), ps.n = v3(tmp.nx, tmp.ny, tmp.nz);
              ps.bits = kHitBit | (tmp.front_face ? kFrontBit : 0u) | (uint32_t)ob->material;
            }
          }
        }
        // cost models for this chunk
        for (int k = 0; k < sc.n_objects; k++) {
          const int ty = sc.objects[k].type;
          const bool cheap = ty == OBJ_SPHERE || ty == OBJ_PLANE || ty == OBJ_QUAD;
          int blk_pass = 0, blk_hit_groups = 0;
          std::vector<int> pass_list;
          for (size_t w0 = 0; w0 < cn; w0 += 32) {
            bool any_reach = false, any_pass = false, any_hit = false, any_pass_hit = false;
            for (size_t j = w0; j < std::min(cn, w0 + 32); j++) {
              const ScanRec &rc = recs[j][(size_t)k];
              any_reach |= rc.reached;
              any_pass |= rc.reached && rc.pass;
              any_hit |= rc.hit;
              any_pass_hit |= rc.hit && rc.pass;
              if (rc.reached) em.tests += 1;
              if (rc.reached && rc.pass) em.tests_pass_cull += 1, pass_list.push_back(rc.hit);
              if (rc.hit) em.hits += 1;
              if (rc.hit && !rc.pass) return -7;  // the cull must be conservative
            }
            if (!any_reach) continue;
            em.warp_obj += 1;
            em.warp_obj_any_pass += any_pass;
            em.now += w.reject[ty] + (any_hit ? w.hit_extra[ty] : 0.0);
            if (cheap) em.lanecull += w.reject[ty] + (any_hit ? w.hit_extra[ty] : 0.0);
            else em.lanecull += w.cull + (any_pass ? w.reject[ty] + (any_hit ? w.hit_extra[ty] : 0.0) : 0.0);
            if (cheap) em.blockcompact += w.reject[ty] + (any_hit ? w.hit_extra[ty] : 0.0);
            else em.blockcompact += w.cull + 2 * w.barrier + 6;
          }
          if (!cheap) {
            blk_pass = (int)pass_list.size();
            for (int g = 0; g < blk_pass; g += 32) {
              bool any_hit = false;
              for (int j = g; j < std::min(blk_pass, g + 32); j++) any_hit |= pass_list[(size_t)j] != 0;
              em.blockcompact += w.reject[ty] + (any_hit ? w.hit_extra[ty] : 0.0) + 10;
              blk_hit_groups += any_hit;
            }
          }
          (void)blk_hit_groups;
        }
        for (size_t w0 = 0; w0 < cn; w0 += 32) em.lanecull += w.cull_setup, em.blockcompact += w.cull_setup;
        {  // (ray, object) pairs for the expensive objects that pass the up-front cull, evaluated densely with
           // t_max = inf, combined per ray in list order (order-independence of the closest-hit scan)
          int np = 0, np_hit = 0;
          std::vector<int> phit;
          for (int k = 0; k < sc.n_objects; k++) {
            const int ty = sc.objects[k].type;
            if (ty != OBJ_CUBE) continue;  // meshes: the object ray moves to the traversal stage's fetch
            for (size_t j = 0; j < cn; j++)
              if (recs[j][(size_t)k].reached && recs[j][(size_t)k].pass_inf) phit.push_back(recs[j][(size_t)k].hit), np++;
          }
          (void)np_hit;
          em.n_pairs += np;
          for (int g = 0; g < np; g += 32) {
            bool any_hit = false;
            for (int j = g; j < std::min(np, g + 32); j++) any_hit |= phit[(size_t)j] != 0;
            em.pairs += w.reject[OBJ_CUBE] + (any_hit ? w.hit_extra[OBJ_CUBE] : 0.0) + 16;  // + load ray/object, store record
          }
          for (size_t w0 = 0; w0 < cn; w0 += 32) {
            em.pairs += w.cull_setup + 3 * w.barrier + 10;
            for (int k = 0; k < sc.n_objects; k++) {
              const int ty = sc.objects[k].type;
              bool any_reach = false, any_hit = false, any_pass = false;
              for (size_t j = w0; j < std::min(cn, w0 + 32); j++) {
                any_reach |= recs[j][(size_t)k].reached, any_hit |= recs[j][(size_t)k].hit;
                any_pass |= recs[j][(size_t)k].reached && recs[j][(size_t)k].pass_inf;
              }
              if (!any_reach) continue;
              if (ty == OBJ_CUBE) em.pairs += w.cull + 8 + (any_pass ? 14 : 0) + (any_hit ? 20 : 0);  // cull + enqueue; combine: filter, copy
              else if (ty == OBJ_MESH) em.pairs += w.cull + (any_pass ? 40 : 0);  // conservative root test only
              else em.pairs += w.reject[ty] + (any_hit ? w.hit_extra[ty] : 0.0);
            }
          }
        }
        {  // sort the chunk by the up-front cull mask (t_max = inf) of the expensive objects, then scan in that order
          std::vector<std::pair<uint32_t, uint32_t>> order;
          for (size_t j = 0; j < cn; j++) {
            uint32_t mask = 0;
            for (int k = 0; k < sc.n_objects; k++) {
              const int ty = sc.objects[k].type;
              if ((ty == OBJ_CUBE || ty == OBJ_MESH) && recs[j][(size_t)k].pass_inf) mask |= 1u << k;
            }
            order.push_back({mask, (uint32_t)j});
          }
          std::stable_sort(order.begin(), order.end(), [](auto &a, auto &b) { return a.first < b.first; });
          for (size_t w0 = 0; w0 < cn; w0 += 32) {
            em.masksort += w.cull_setup + 60 + 3 * w.barrier;  // masks are computed below; sort + gather overhead
            for (int k = 0; k < sc.n_objects; k++) {
              const int ty = sc.objects[k].type;
              const bool cheap = ty == OBJ_SPHERE || ty == OBJ_PLANE || ty == OBJ_QUAD;
              bool any = false, any_hit = false, any_reach = false;
              for (size_t j = w0; j < std::min(cn, w0 + 32); j++) {
                const ScanRec &rc = recs[order[j].second][(size_t)k];
                any_reach |= rc.reached;
                any |= rc.reached && (cheap || rc.pass_inf);
                any_hit |= rc.hit;
              }
              if (!any_reach) continue;
              if (!cheap) em.masksort += w.cull;
              if (any) em.masksort += w.reject[ty] + (any_hit ? w.hit_extra[ty] : 0.0);
            }
          }
        }
      }
      (void)first_pass;
      first_pass = false;
      if (parked.empty()) break;
      n_tasks += (double)parked.size();
      // ---- traverse under every policy (results are policy-independent; the last run's results are used)
      for (int p = 0; p < mp->n_policies; p++) {
        TravPolicy pol;
        pol.refill_lanes = mp->policy[p][0], pol.mode = mp->policy[p][1], pol.thr_node = mp->policy[p][2], pol.thr_tri = mp->policy[p][3];
        pol.tri_groups = mp->policy[p][4];
        std::vector<Task> copy = parked;
        if (pol.mode >= 10) {  // tasks sorted by a spatial key before the traversal (v1 stepping policy)
          const int variant = pol.mode;
          pol.mode = 0;
          std::vector<std::pair<uint32_t, uint32_t>> keys(copy.size());
          for (size_t i = 0; i < copy.size(); i++) {
            const Task &t = copy[i];
            const DMesh &gm = sc.meshes[sc.objects[t.object].mesh];
            const uint32_t oct = (t.d.x < 0 ? 1u : 0u) | (t.d.y < 0 ? 2u : 0u) | (t.d.z < 0 ? 4u : 0u);
            uint32_t key = oct;
            if (variant >= 11) {
              // entry point of the ray into the root frame (variant 11) or its origin (12), 4 bits per axis, Morton order
              float tn = 0.0f;
              if (variant == 11) {
                const float o[3] = {t.o.x, t.o.y, t.o.z}, d[3] = {t.d.x, t.d.y, t.d.z};
                for (int a = 0; a < 3; a++) {
                  const float inv = 1.0f / d[a];
                  const float t0 = (gm.root_lo[a] - o[a]) * inv, t1 = (gm.root_hi[a] - o[a]) * inv;
                  tn = fmaxf(tn, fminf(t0, t1));
                }
              }
              const float pnt[3] = {t.o.x + t.d.x * tn, t.o.y + t.d.y * tn, t.o.z + t.d.z * tn};
              uint32_t m = 0;
              for (int a = 0; a < 3; a++) {
                float u = (pnt[a] - gm.root_lo[a]) / std::max(gm.root_hi[a] - gm.root_lo[a], 1e-20f);
                u = std::min(std::max(u, 0.0f), 0.999f);
                const uint32_t q = (uint32_t)(u * 16.0f);
                for (int b = 0; b < 4; b++) m |= ((q >> b) & 1u) << (3 * b + a);
              }
              key = (oct << 12) | m;
            }
            keys[i] = {key, (uint32_t)i};
          }
          std::stable_sort(keys.begin(), keys.end(), [](auto &a, auto &b) { return a.first < b.first; });
          std::vector<Task> sorted(copy.size());
          for (size_t i = 0; i < copy.size(); i++) sorted[i] = copy[keys[i].second];
          run_traverse(sc, sorted, t_min, pol, w, tm[p]);
          for (size_t i = 0; i < copy.size(); i++) copy[keys[i].second] = sorted[i];
        } else {
          run_traverse(sc, copy, t_min, pol, w, tm[p]);
        }
        if (p == mp->n_policies - 1) parked.swap(copy);
      }
      // ---- post: finish + rest of the list
      todo.clear();
      for (const Task &t : parked) {
        PathState &ps = pool[t.ray];
        const Ray ray{ps.o, ps.d};
        float &cl = closest[t.ray];
        if (t.res_tri != 0xffffffffu) {
          const DObject *ob = sc.objects + t.object;
          const MeshRay omr = mesh_object_ray(ob->f, ray);
          MeshHit mh;
          mh.t = t.res_t, mh.tri = t.res_tri, mh.order = 0u;
          Hit tmp;
          if (mesh_finish(ob->f, sc.meshes[ob->mesh], ray, omr, mh, t_min, cl, tmp)) {
            cl = tmp.t;
            ps.t = tmp.t, ps.p = v3(tmp.px, tmp.py, tmp.pz), ps.n = v3(tmp.nx, tmp.ny, tmp.nz);
            ps.bits = kHitBit | (tmp.front_face ? kFrontBit : 0u) | (uint32_t)ob->material;
          }
        }
        todo.push_back(Pending{t.ray, t.object + 1});
      }
    }
    {  // shade imbalance model: class weights (warp instructions per homogeneous 32-ray group)
      const double cw[10] = {40, 300, 320, 330, 60, 260, 380, 650, 40, 0};  // miss, lambert, checker, metal, emissive?, dielectric, plastic, conductor, null
      for (int v = 0; v < 2; v++) {
        const uint32_t win = v == 0 ? 256u : 2048u;
        for (uint32_t c0 = 0; c0 < n; c0 += win) {
          const uint32_t cn = std::min<uint32_t>(win, n - c0);
          std::vector<uint32_t> keys;
          for (uint32_t j = 0; j < cn; j++) {
            const uint32_t bits = pool[c0 + j].bits;
            keys.push_back((bits & kHitBit) ? 1u + (uint32_t)sc.materials[bits & kMatMask].type : 0u);
          }
          std::sort(keys.begin(), keys.end());
          double per_warp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          uint32_t g = 0;
          for (uint32_t w0 = 0; w0 < cn; w0 += 32, g++) {
            uint32_t classes = 0;
            for (uint32_t j = w0; j < std::min(cn, w0 + 32); j++) classes |= 1u << keys[j];
            double c = 0;
            for (int k = 0; k < 10; k++)
              if (classes >> k & 1u) c += cw[k];
            per_warp[g % 8] += c;
            sh_sum[v] += c;
          }
          sh_sync[v] += 8.0 * *std::max_element(per_warp, per_warp + 8);
        }
      }
    }
    // ---- shade: chunks of 256, stable counting sort by class, survivors in sorted order
    uint32_t wcur = 0;
    std::vector<PathState> next;
    next.reserve(n);
    for (uint32_t c0 = 0; c0 < n; c0 += kBlock) {
      const uint32_t cn = std::min<uint32_t>(kBlock, n - c0);
      std::vector<std::pair<uint32_t, uint32_t>> order;  // (class, index)
      for (uint32_t j = 0; j < cn; j++) {
        const uint32_t bits = pool[c0 + j].bits;
        const uint32_t key = (bits & kHitBit) ? 1u + (uint32_t)sc.materials[bits & kMatMask].type : 0u;
        order.push_back({key, j});
      }
      std::stable_sort(order.begin(), order.end(), [](auto &a, auto &b) { return a.first < b.first; });
      for (uint32_t w0 = 0; w0 < cn; w0 += 32) {  // shade divergence: distinct classes per warp
        uint32_t classes = 0;
        for (uint32_t j = w0; j < std::min(cn, w0 + 32); j++) classes |= 1u << order[j].first;
        shade_lane_eff_num += std::min(cn, w0 + 32) - w0;
        shade_lane_eff_den += 32.0 * popc32(classes);
      }
      for (uint32_t j = 0; j < cn; j++) {
        const PathState &ps = pool[c0 + order[j].second];
        if (!(ps.bits & kHitBit)) continue;  // sky: terminates
        const DMaterial m = sc.materials[ps.bits & kMatMask];
        const Uniforms4 u = philox_uniforms(mp->seed, ps.pixel, ps.sample, ps.bounce, 0u);
        Ray sr;
        V3 att;
        if (mat_scatter(m, ps.d, ps.p, ps.n, (ps.bits & kFrontBit) != 0u, u.u, sr, att)) {
          if (ps.bounce + 1u < (uint32_t)mp->max_depth) {
            PathState q = ps;
            q.o = sr.o, q.d = sr.d, q.beta = ps.beta * att, q.bounce = ps.bounce + 1u;
            next.push_back(q);
          }
        }
      }
    }
    wcur = (uint32_t)next.size();
    for (uint32_t i = 0; i < wcur; i++) pool[i] = next[i];
    n = wcur;
    regenerate();
  }

  int o = 0;
  auto put = [&](double v) {
    if (o < n_out) out[o] = v;
    o++;
  };
  put(rays), put(n_tasks);
  put(em.pairs), put(em.n_pairs);
  put(em.masksort);
  put(em.now), put(em.lanecull), put(em.blockcompact), put(em.tests), put(em.tests_pass_cull), put(em.hits), put(em.warp_obj),
      put(em.warp_obj_any_pass);
  put(shade_lane_eff_num / std::max(1.0, shade_lane_eff_den));
  put(sh_sum[0]), put(sh_sync[0]), put(sh_sum[1]), put(sh_sync[1]);
  for (int p = 0; p < mp->n_policies; p++) {
    const TravModel &t = tm[p];
    put(t.cost), put(t.ideal), put(t.turns), put(t.node_exec), put(t.tri_exec), put(t.node_lanes), put(t.tri_lanes), put(t.refills);
  }
  for (int i = 0; i < 16; i++) put(tm[0].hist_nodes[i]);
  for (int i = 0; i < 16; i++) put(tm[0].hist_tris[i]);
  put(tm[0].hits);
  return o;
}

}  // extern "C"
