// ptcore.cu — the B200 path-tracing core: wavefront kernels for sm_100a + the C ABI of include/ptcore.h.
//
// Replaces the body of render_scene (src/renderer.rs:87-106 of the reference) and everything trace_ray
// (renderer.rs:19-65) calls.  trace_ray's recursion  L = Le + f * trace(next)  is unrolled into a wavefront:
//
//   generate : path index -> (pixel, sample) -> Philox jitter -> Camera::get_ray          (renderer.rs:96-99)
//   extend   : closest hit of every live ray (HittableList::hit + primitives + BVH)        (renderer.rs:24)
//   shade    : miss -> sky; hit -> emitted + Material::scatter; throughput update;         (renderer.rs:26-63)
//              survivors are compacted into the other ray buffer (warp ballot + one atomic per warp)
//   advance  : one thread; tops the survivor buffer up with fresh camera paths (path regeneration) so the
//              wavefront stays full although most paths leave after a few segments
//
// Only Emissive surfaces and the sky carry radiance and both end the path (EmissiveLight::scatter is None), so a
// path contributes beta * Le exactly once, when it terminates: one float RED triple per path into the film.
// extend runs persistent warps that fetch 32 rays at a time from a device-side counter; every kernel reads its
// item count from device memory, so the host enqueues iterations without synchronising and only polls a
// pinned copy of the control block every few iterations.
//
// There is no CPU fallback in this library: without a CUDA device commit / render / intersect return PTC_E_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <string>
#include <vector>

#include "../../include/ptcore.h"
#include "pt_bsdf.h"
#include "pt_philox.h"
#include "pt_prims.h"
#include "pt_scene_host.h"

using namespace pt;

namespace {

thread_local std::string g_err;

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

// ------------------------------------------------------------------------------------------------------------
// Device-side control block of one render (lives in device memory, mirrored to pinned host memory for polling)
struct Ctl {
  unsigned long long next_path;   // next camera path index to start
  unsigned long long total_paths; // path index space of this call (padded tiles x samples)
  unsigned long long rays;        // extend items so far = trace_ray calls with depth > 0
  unsigned long long nodes, tris, mesh_rays;  // PTC_FLAG_COUNTERS
  uint32_t n_cur;        // rays in the current buffer
  uint32_t n_next;       // survivors appended to the other buffer by shade
  uint32_t work_extend;  // dynamic fetch cursor of the persistent extend warps
  uint32_t gen_count;    // camera paths generate has to start this iteration
  unsigned long long gen_first_path;
  uint32_t iterations;
  uint32_t done;
};

// Per extend round r (= r-th mesh object a ray is parked at): number of tasks, and the dynamic fetch cursor of the
// traversal warps.  Sized by the number of mesh objects of the scene (+1), zeroed by k_advance.
struct RoundCtl {
  uint32_t *n;
  uint32_t *cursor;
  int32_t rounds;
};

struct RenderParams {
  DCamera cam;
  int32_t width, height;
  int32_t max_depth;
  int32_t sample_begin, n_samples;
  int32_t tiles_x, n_my_tiles, tile_mod, tile_rem;
  uint32_t pool;
  uint64_t seed;
};

struct Buffers {
  float4 *ray_o[2];  // origin.xyz, pixel index
  float4 *ray_d[2];  // direction.xyz, sample index
  float4 *beta[2];   // throughput.rgb, bounce (segments already traced)
  float4 *hit0;      // position.xyz, t
  float4 *hit1;      // normal.xyz, [hit<<31 | front_face<<30 | material]
};

constexpr uint32_t kHitBit = 0x80000000u, kFrontBit = 0x40000000u, kMatMask = 0x3fffffffu;

// ------------------------------------------------------------------------------------------------------------
__global__ void k_advance(Ctl *ctl, RoundCtl rc, uint32_t pool) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (ctl->n_cur != 0) {
    ctl->rays += ctl->n_cur;
    ctl->iterations++;
  }
  const uint32_t n = ctl->n_next;
  const unsigned long long remaining = ctl->total_paths - ctl->next_path;
  const unsigned long long space = pool - n;
  const uint32_t g = (uint32_t)(remaining < space ? remaining : space);
  ctl->gen_first_path = ctl->next_path;
  ctl->gen_count = g;
  ctl->next_path += g;
  ctl->n_cur = n;  // generate appends the valid camera rays behind the survivors
  ctl->n_next = 0;
  ctl->work_extend = 0;
  for (int r = 0; r <= rc.rounds; r++) rc.n[r] = 0u, rc.cursor[r] = 0u;
  ctl->done = (g == 0 && n == 0) ? 1u : 0u;
}

// path index -> pixel/sample.  Paths are ordered sample-major, then by this rank's 32x32 tiles, then row-major inside
// a tile, so a warp starts 32 horizontally adjacent pixels of one sample (coherent primary rays, distinct film
// addresses).  Pixels of partial border tiles that fall outside the image start no path.
__global__ void __launch_bounds__(256) k_generate(Ctl *ctl, RenderParams rp, Buffers b, int dst) {
  const uint32_t count = ctl->gen_count;
  const unsigned long long first = ctl->gen_first_path;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long per_sample = (unsigned long long)rp.n_my_tiles * 1024ull;
  for (uint32_t base = warp_id * 32u; base < count; base += warps_total * 32u) {
    const uint32_t j = base + lane;
    bool valid = j < count;
    uint32_t pixel = 0, sample = 0;
    int x = 0, y = 0;
    if (valid) {
      const unsigned long long p = first + j;
      const uint32_t s = (uint32_t)(p / per_sample);
      const uint32_t r = (uint32_t)(p - (unsigned long long)s * per_sample);
      const uint32_t local_tile = r >> 10, in_tile = r & 1023u;
      const uint32_t tile = local_tile * (uint32_t)rp.tile_mod + (uint32_t)rp.tile_rem;
      x = (int)((tile % (uint32_t)rp.tiles_x) * 32u + (in_tile & 31u));
      y = (int)((tile / (uint32_t)rp.tiles_x) * 32u + (in_tile >> 5));
      valid = x < rp.width && y < rp.height;
      pixel = (uint32_t)(y * rp.width + x);
      sample = (uint32_t)rp.sample_begin + s;
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, valid);
    if (mask == 0u) continue;
    uint32_t slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(&ctl->n_cur, (uint32_t)__popc(mask));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (valid) {
      const uint32_t slot = slot0 + (uint32_t)__popc(mask & ((1u << lane) - 1u));
      const Uniforms4 jit = philox_uniforms(rp.seed, pixel, sample, 0xffffffffu, 0u);
      const float u = ((float)x + jit.u[0]) / (float)rp.width;   // renderer.rs:96
      const float v = ((float)y + jit.u[1]) / (float)rp.height;  // renderer.rs:97
      const Ray ray = camera_get_ray(rp.cam, u, v);
      b.ray_o[dst][slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, u2f(pixel));
      b.ray_d[dst][slot] = make_float4(ray.d.x, ray.d.y, ray.d.z, u2f(sample));
      b.beta[dst][slot] = make_float4(1.0f, 1.0f, 1.0f, u2f(0u));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// extend = HittableList::hit (hittable.rs:46-57) for every ray of the current buffer, as three kernels.
//
// The object list is scanned in insertion order exactly like the reference, but a thread never walks a BVH in line
// with the analytic primitives.  Measured on the first version (one thread = one ray = the whole list, profiles/r1_v1_*):
// the traversal code ran with 7.5 of 32 lanes active and the triangle test with 2, because most rays miss a mesh's
// root box and the rest need wildly different numbers of steps.  So:
//
//   k_extend_pre   every ray, full warps: analytic primitives up to the first mesh whose (conservative) root-frame test
//                  passes; such a ray is PARKED: its partial closest hit goes to the hit record it owns, and a task
//                  {ray, object index, object-space ray, closest_so_far} is appended to a queue (one atomic per warp)
//   k_traverse     persistent warps over the task queue: each lane walks its own BVH one step at a time (one node or
//                  one triangle per turn) and a lane that finishes fetches the next task at once, so warps stay
//                  populated however uneven the rays are
//   k_extend_post  every task, full warps: second half of Mesh::hit (world-space record, closed-interval recheck), then
//                  the scan resumes at the next object; a ray that meets another mesh is parked again (next round)
//
// Every object still sees exactly the `closest_so_far` it would have seen in the reference's sequential scan, so the
// tie-breaking and interval conventions of hittable.rs:50-55 are untouched.
constexpr int kExtendThreads = 128;
constexpr uint32_t kRefillLanes = 8;  // a traversal warp fetches new tasks once this many lanes are idle

struct ExtendOut {
  Buffers b;
  int2 *ids;  // (object, triangle) per ray, only for ptc_intersect (nullptr in renders)
};

struct TaskQ {       // [round & 1]
  uint2 *ray[2];     // x = ray index, y = object index of the mesh
  float4 *o[2];      // object-space origin, closest_so_far (the t_max Mesh::hit was called with)
  float4 *d[2];      // object-space direction (normalised twice, mesh_object.rs:287 + ray.rs:15)
  float2 *res[2];    // traversal result: t (object space), original triangle index or 0xffffffff
};

__device__ __forceinline__ void write_hit(const ExtendOut &out, uint32_t i, const Hit &h) {
  out.b.hit0[i] = make_float4(h.px, h.py, h.pz, h.t);
  out.b.hit1[i] = make_float4(h.nx, h.ny, h.nz, u2f(kHitBit | (h.front_face ? kFrontBit : 0u) | ((uint32_t)h.material & kMatMask)));
  if (out.ids) out.ids[i] = make_int2(h.object, h.triangle);
}
__device__ __forceinline__ void write_miss(const ExtendOut &out, uint32_t i) {
  out.b.hit1[i] = make_float4(0.0f, 0.0f, 0.0f, u2f(0u));
  if (out.ids) out.ids[i] = make_int2(-1, -1);
}

// Scan objects [k_begin, n).  Returns the index of the mesh the ray has to be parked at (its object-space ray in
// `park_ray`), or -1 when the scan is complete.
__device__ __forceinline__ int scan_objects(const DScene &sc, const Ray &ray, float t_min, float &closest, Hit &best,
                                            bool &improved, int k_begin, MeshRay &park_ray) {
  for (int k = k_begin; k < sc.n_objects; k++) {
    const DObject *ob = sc.objects + k;
    const int type = ob->type;
    Hit tmp;
    tmp.triangle = -1;
    bool hit;
    if (type == OBJ_MESH) {
      const DMesh &mesh = sc.meshes[ob->mesh];
      const MeshRay mr = mesh_object_ray(ob->f, ray);
      if (!mesh_root_may_hit(mesh, mr, t_min, closest)) continue;  // cannot hit: same as Mesh::hit returning None
      park_ray = mr;
      return k;
    }
    if (type == OBJ_SPHERE) hit = hit_sphere(ob->f, ray, t_min, closest, tmp);
    else if (type == OBJ_QUAD) hit = hit_quad(ob->f, ray, t_min, closest, tmp);
    else if (type == OBJ_CUBE) hit = hit_cube(ob->f, ray, t_min, closest, tmp);
    else hit = hit_plane(ob->f, ray, t_min, closest, tmp);
    if (hit) {
      improved = true;
      closest = tmp.t;
      best = tmp;
      best.object = k;
      best.material = ob->material;
    }
  }
  return -1;
}

// warp-aggregated append of the parked lanes to the task queue of round `round`
__device__ __forceinline__ void park_tasks(const RoundCtl &rc, const TaskQ &tq, int round, int park, uint32_t i, const MeshRay &mr,
                                           float closest) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t mask = __ballot_sync(0xffffffffu, park >= 0);
  if (mask == 0u) return;
  uint32_t slot0 = 0;
  if (lane == 0) slot0 = atomicAdd(&rc.n[round], (uint32_t)__popc(mask));
  slot0 = __shfl_sync(0xffffffffu, slot0, 0);
  if (park >= 0) {
    const uint32_t slot = slot0 + (uint32_t)__popc(mask & ((1u << lane) - 1u));
    const int par = round & 1;
    tq.ray[par][slot] = make_uint2(i, (uint32_t)park);
    tq.o[par][slot] = make_float4(mr.o.x, mr.o.y, mr.o.z, closest);
    tq.d[par][slot] = make_float4(mr.d.x, mr.d.y, mr.d.z, 0.0f);
  }
}

template <bool COUNT>
__global__ void __launch_bounds__(kExtendThreads) k_extend_pre(Ctl *ctl, RoundCtl rc, DScene sc, ExtendOut out, TaskQ tq, int src,
                                                               float t_min, float t_max) {
  const uint32_t n = ctl->n_cur;
  const uint32_t lane = threadIdx.x & 31u;
  TraversalCounters tc{0u, 0u, 0u};
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&ctl->work_extend, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n) break;
    const uint32_t i = base + lane;
    int park = -1;
    MeshRay mr;
    float closest = t_max;
    if (i < n) {
      const float4 o4 = out.b.ray_o[src][i], d4 = out.b.ray_d[src][i];
      const Ray ray{v3(o4.x, o4.y, o4.z), v3(d4.x, d4.y, d4.z)};
      Hit best;
      bool improved = false;
      park = scan_objects(sc, ray, t_min, closest, best, improved, 0, mr);  // renderer.rs:24
      if (improved) write_hit(out, i, best);
      else write_miss(out, i);
    }
    park_tasks(rc, tq, 0, park, i, mr, closest);
  }
  if (COUNT) {
    atomicAdd(&ctl->nodes, (unsigned long long)tc.nodes);
    atomicAdd(&ctl->tris, (unsigned long long)tc.tris);
    atomicAdd(&ctl->mesh_rays, (unsigned long long)tc.mesh_rays);
  }
}

template <bool COUNT>
__global__ void __launch_bounds__(kExtendThreads) k_traverse(Ctl *ctl, RoundCtl rc, DScene sc, TaskQ tq, int round, float t_min, uint32_t refill_lanes) {
  const uint32_t n = rc.n[round];
  const int par = round & 1;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  TraversalCounters tc{0u, 0u, 0u};
  TravState s;
  DMesh mesh;  // the two pointers of the mesh this lane walks, kept in registers
  mesh.nodes = nullptr, mesh.tris = nullptr;
  uint2 stack[kTraversalStack];
  int sp = 0;
  bool active = false;
  uint32_t task = 0;
  bool exhausted = n == 0u;
  for (;;) {
    const uint32_t idle = __ballot_sync(0xffffffffu, !active);
    if (!exhausted && (idle == 0xffffffffu || (uint32_t)__popc(idle) >= refill_lanes)) {
      const uint32_t cnt = (uint32_t)__popc(idle);
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&rc.cursor[round], cnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (!active) {
        const uint32_t j = base + (uint32_t)__popc(idle & lt_mask);
        if (j < n) {
          const uint2 rk = tq.ray[par][j];
          const float4 o4 = tq.o[par][j], d4 = tq.d[par][j];
          const DMesh *gm = sc.meshes + sc.objects[rk.y].mesh;
          mesh.nodes = gm->nodes;
          mesh.tris = gm->tris;
          trav_begin(s, v3(o4.x, o4.y, o4.z), v3(d4.x, d4.y, d4.z), t_min, o4.w);  // world t bounds, mesh_object.rs:289-291
          sp = 0;
          task = j;
          active = true;
          if (COUNT) tc.mesh_rays++;
        }
      }
      if (base + cnt >= n) exhausted = true;
    }
    if (__ballot_sync(0xffffffffu, active) == 0u) {
      if (exhausted) break;
      continue;
    }
    if (active) {
      if (!trav_has_tri(s) && !trav_has_node(s)) {
        if (sp > 0) {
          s.ng = stack[--sp];
        } else {
          tq.res[par][task] = make_float2(s.best_t, u2f(s.best_tri));
          active = false;
        }
      }
      if (active && !trav_has_tri(s) && trav_has_node(s)) trav_node<COUNT>(mesh, s, stack, sp, &tc);
      if (active && trav_has_tri(s)) trav_tri<COUNT>(mesh, s, &tc);
    }
  }
  if (COUNT) {
    atomicAdd(&ctl->nodes, (unsigned long long)tc.nodes);
    atomicAdd(&ctl->tris, (unsigned long long)tc.tris);
    atomicAdd(&ctl->mesh_rays, (unsigned long long)tc.mesh_rays);
  }
}

template <bool COUNT>
__global__ void __launch_bounds__(kExtendThreads) k_extend_post(Ctl *ctl, RoundCtl rc, DScene sc, ExtendOut out, TaskQ tq, int src,
                                                                int round, float t_min, float t_max) {
  const uint32_t n = rc.n[round];
  const int par = round & 1;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  TraversalCounters tc{0u, 0u, 0u};
  for (uint32_t base = warp_id * 32u; base < n; base += warps_total * 32u) {
    const uint32_t j = base + lane;
    int park = -1;
    MeshRay mr;
    float closest = t_max;
    uint32_t i = 0;
    if (j < n) {
      const uint2 rk = tq.ray[par][j];
      const float2 res = tq.res[par][j];
      i = rk.x;
      const int k = (int)rk.y;
      const float4 o4 = out.b.ray_o[src][i], d4 = out.b.ray_d[src][i];
      const Ray ray{v3(o4.x, o4.y, o4.z), v3(d4.x, d4.y, d4.z)};
      const bool any = (f2u(out.b.hit1[i].w) & kHitBit) != 0u;
      if (any) closest = out.b.hit0[i].w;  // == the t_max the traversal ran with
      Hit best;
      best.triangle = -1;
      bool improved = false;
      const uint32_t tri = f2u(res.y);
      if (tri != 0xffffffffu) {
        const DObject *ob = sc.objects + k;
        const MeshRay omr = mesh_object_ray(ob->f, ray);  // same inputs, same bits as in k_extend_pre
        MeshHit mh;
        mh.t = res.x, mh.tri = tri, mh.order = 0u;
        Hit tmp;
        if (mesh_finish(ob->f, sc.meshes[ob->mesh], ray, omr, mh, t_min, closest, tmp)) {
          improved = true;
          closest = tmp.t;
          best = tmp;
          best.object = k;
          best.material = ob->material;
        }
      }
      park = scan_objects(sc, ray, t_min, closest, best, improved, k + 1, mr);
      if (improved) write_hit(out, i, best);  // otherwise the record parked by the previous stage stands
    }
    park_tasks(rc, tq, round + 1, park, i, mr, closest);
  }
  if (COUNT) {
    atomicAdd(&ctl->nodes, (unsigned long long)tc.nodes);
    atomicAdd(&ctl->tris, (unsigned long long)tc.tris);
    atomicAdd(&ctl->mesh_rays, (unsigned long long)tc.mesh_rays);
  }
}

constexpr int kShadeThreads = 256;
constexpr int kShadeClasses = 10;  // 0 = miss, 1 + material type (8 types), 9 = no ray (tail of the last chunk)
constexpr int kShadeArrays = 5;    // ray_o, ray_d, beta, hit0, hit1
constexpr size_t kShadeSmem = 2 * kShadeArrays * kShadeThreads * sizeof(float4) + 128;  // two staged chunks + alignment slack

// ---- sm_100a asynchronous bulk copy (TMA, 1-D) + mbarrier, raw PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared, completion (byte count) signalled on the mbarrier; SASS: UBLKCP
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shade: emitted + scatter (renderer.rs:26-36) or sky (renderer.rs:38-63); compacts survivors into buffer `dst`.
//
// Rays arrive in no particular order, so a warp would see a mix of misses and of every material and run all of their
// code (measured: 11.5 of 32 lanes active), and the kernel is latency-bound if each thread waits for its own loads.
// Each persistent block therefore walks chunks of 256 consecutive rays with a two-stage pipeline:
//   * one thread issues five 1-D bulk async copies (TMA; ray_o, ray_d, beta, hit0, hit1 slices, 4 KB each) of the NEXT
//     chunk into shared memory, completion counted on an mbarrier, while the block shades the current one;
//   * the current chunk is counting-sorted by (miss | material type) in shared memory (10 ballots per warp + one
//     10-lane scan); thread t then shades the t-th ray of that order straight out of the staged copy, so warps are
//     homogeneous except where a class boundary falls inside them (26 of 32 lanes active after, 11.5 before);
//   * survivors are compacted into the other ray buffer with one atomic per warp.
__global__ void __launch_bounds__(kShadeThreads, 3) k_shade(Ctl *ctl, DScene sc, RenderParams rp, Buffers b, int src, int dst, float *accum) {
  extern __shared__ uint8_t s_dyn[];
  float4 *s_raw = reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(s_dyn) + 127) & ~(uintptr_t)127);  // [2][5][256]
  __shared__ uint16_t s_perm[kShadeThreads];
  __shared__ uint32_t s_cnt[kShadeThreads / 32][kShadeClasses];
  __shared__ uint32_t s_off[kShadeThreads / 32][kShadeClasses];
  __shared__ __align__(8) uint64_t s_bar[2];
  const uint32_t n = ctl->n_cur;
  const uint32_t n_chunks = (n + (uint32_t)kShadeThreads - 1) / (uint32_t)kShadeThreads;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const float4 *g_arr[kShadeArrays] = {b.ray_o[src], b.ray_d[src], b.beta[src], b.hit0, b.hit1};

  auto issue = [&](uint32_t chunk, uint32_t buf) {  // one thread
    const uint32_t first = chunk * (uint32_t)kShadeThreads;
    const uint32_t cnt = n - first < (uint32_t)kShadeThreads ? n - first : (uint32_t)kShadeThreads;
    const uint32_t bytes = cnt * (uint32_t)sizeof(float4);
    mbar_expect_tx(&s_bar[buf], bytes * kShadeArrays);
#pragma unroll
    for (int a = 0; a < kShadeArrays; a++)
      bulk_g2s(s_raw + ((size_t)buf * kShadeArrays + a) * kShadeThreads, g_arr[a] + first, bytes, &s_bar[buf]);
  };

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x < n_chunks) issue(blockIdx.x, 0);

  uint32_t k = 0;
  for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, k++) {
    const uint32_t buf = k & 1u;
    // the other stage was last read in the previous iteration, which ended with a block barrier
    if (tid == 0 && chunk + gridDim.x < n_chunks) issue(chunk + gridDim.x, buf ^ 1u);
    const uint32_t first = chunk * (uint32_t)kShadeThreads;
    const uint32_t in_chunk = n - first < (uint32_t)kShadeThreads ? n - first : (uint32_t)kShadeThreads;
    const float4 *raw = s_raw + (size_t)buf * kShadeArrays * kShadeThreads;
    mbar_wait(&s_bar[buf], (k >> 1) & 1u);

    uint32_t key = kShadeClasses - 1;
    if (tid < in_chunk) {
      const uint32_t bits0 = f2u(raw[4 * kShadeThreads + tid].w);
      key = (bits0 & kHitBit) ? 1u + (uint32_t)sc.materials[bits0 & kMatMask].type : 0u;
    }
    uint32_t rank = 0;
#pragma unroll
    for (uint32_t c = 0; c < (uint32_t)kShadeClasses; c++) {
      const uint32_t m = __ballot_sync(0xffffffffu, key == c);
      if (key == c) rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (lane == 0) s_cnt[warp][c] = (uint32_t)__popc(m);
    }
    __syncthreads();
    if (warp == 0) {
      uint32_t tot = 0;
      if (lane < (uint32_t)kShadeClasses)
        for (int w = 0; w < kShadeThreads / 32; w++) tot += s_cnt[w][lane];
      uint32_t incl = tot;  // inclusive scan over the classes
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += v;
      }
      if (lane < (uint32_t)kShadeClasses) {
        uint32_t off = incl - tot;
        for (int w = 0; w < kShadeThreads / 32; w++) {
          s_off[w][lane] = off;
          off += s_cnt[w][lane];
        }
      }
    }
    __syncthreads();
    s_perm[s_off[warp][key] + rank] = (uint16_t)tid;
    __syncthreads();

    bool alive = false;
    float4 no, nd, nb;
    if (tid < in_chunk) {  // the "no ray" class sorts last
      const uint32_t j = s_perm[tid];
      const float4 o4 = raw[0 * kShadeThreads + j], d4 = raw[1 * kShadeThreads + j], b4 = raw[2 * kShadeThreads + j];
      const float4 h1 = raw[4 * kShadeThreads + j];
      const uint32_t pixel = f2u(o4.w), sample = f2u(d4.w), bounce = f2u(b4.w);
      const uint32_t bits = f2u(h1.w);
      const V3 beta = v3(b4.x, b4.y, b4.z);
      const V3 ray_d = v3(d4.x, d4.y, d4.z);
      V3 radiance = v3(0, 0, 0);
      bool add = false;
      if (!(bits & kHitBit)) {
        radiance = beta * sky_color(sc, ray_d);
        add = true;
      } else {
        const float4 h0 = raw[3 * kShadeThreads + j];
        const DMaterial m = sc.materials[bits & kMatMask];
        const V3 e = mat_emitted(m);
        if (e.x != 0.0f || e.y != 0.0f || e.z != 0.0f) {
          radiance = beta * e;
          add = true;
        }
        const Uniforms4 u = philox_uniforms(rp.seed, pixel, sample, bounce, 0u);
        Ray sc_ray;
        V3 att;
        if (mat_scatter(m, ray_d, v3(h0.x, h0.y, h0.z), v3(h1.x, h1.y, h1.z), (bits & kFrontBit) != 0u, u.u, sc_ray, att)) {
          // trace_ray(scattered, depth - 1): depth 0 returns black (renderer.rs:20-22)
          if (bounce + 1u < (uint32_t)rp.max_depth) {
            alive = true;
            const V3 nbeta = beta * att;
            no = make_float4(sc_ray.o.x, sc_ray.o.y, sc_ray.o.z, o4.w);
            nd = make_float4(sc_ray.d.x, sc_ray.d.y, sc_ray.d.z, d4.w);
            nb = make_float4(nbeta.x, nbeta.y, nbeta.z, u2f(bounce + 1u));
          }
        }
      }
      if (add) {
        float *px = accum + (size_t)pixel * 3;
        atomicAdd(px + 0, radiance.x);
        atomicAdd(px + 1, radiance.y);
        atomicAdd(px + 2, radiance.z);
      }
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, alive);
    if (mask != 0u) {
      uint32_t slot0 = 0;
      if (lane == 0) slot0 = atomicAdd(&ctl->n_next, (uint32_t)__popc(mask));
      slot0 = __shfl_sync(0xffffffffu, slot0, 0);
      if (alive) {
        const uint32_t slot = slot0 + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        b.ray_o[dst][slot] = no;
        b.ray_d[dst][slot] = nd;
        b.beta[dst][slot] = nb;
      }
    }
    __syncthreads();  // stage `buf` and s_perm are free again
  }
}

// out = rgb * scale (renderer.rs:103)
__global__ void k_scale(const float *in, float *out, size_t n, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * scale;
}
// renderer.rs:112-120 + color.rs:87-93
__global__ void k_resolve(const float *rgb, size_t n_pixels, float scale, uint32_t *out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pixels) out[i] = resolve_pixel(rgb[i * 3] * scale, rgb[i * 3 + 1] * scale, rgb[i * 3 + 2] * scale);
}

// ---- parity hooks: the same device functions, driven by caller-provided inputs ------------------------------
// ptc_intersect runs the very kernel the renderer uses (k_extend); these two only repack its inputs / outputs.
__global__ void k_pack_rays(const float *o, const float *d, size_t n, float4 *ro, float4 *rd) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ro[i] = make_float4(o[i * 3], o[i * 3 + 1], o[i * 3 + 2], 0.0f);
  rd[i] = make_float4(d[i * 3], d[i * 3 + 1], d[i * 3 + 2], 0.0f);
}
__global__ void k_unpack_hits(const float4 *hit0, const float4 *hit1, const int2 *ids, size_t n, ptc_hit *out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ptc_hit r;
  memset(&r, 0, sizeof(r));
  const float4 h1 = hit1[i];
  const uint32_t bits = f2u(h1.w);
  if (bits & kHitBit) {
    const float4 h0 = hit0[i];
    const int2 id = ids[i];
    r.object = id.x, r.triangle = id.y;
    r.t = h0.w;
    r.position[0] = h0.x, r.position[1] = h0.y, r.position[2] = h0.z;
    r.normal[0] = h1.x, r.normal[1] = h1.y, r.normal[2] = h1.z;
    r.front_face = (bits & kFrontBit) ? 1 : 0;
    r.material = (int32_t)(bits & kMatMask);
  } else {
    r.object = -1, r.triangle = -1, r.material = -1;
  }
  out[i] = r;
}

__global__ void k_primary_rays(RenderParams rp, uint32_t sample, float *out_o, float *out_d) {
  const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
  if (pixel >= (uint32_t)(rp.width * rp.height)) return;
  const int x = (int)(pixel % (uint32_t)rp.width), y = (int)(pixel / (uint32_t)rp.width);
  const Uniforms4 jit = philox_uniforms(rp.seed, pixel, sample, 0xffffffffu, 0u);
  const float u = ((float)x + jit.u[0]) / (float)rp.width;
  const float v = ((float)y + jit.u[1]) / (float)rp.height;
  const Ray ray = camera_get_ray(rp.cam, u, v);
  out_o[pixel * 3 + 0] = ray.o.x, out_o[pixel * 3 + 1] = ray.o.y, out_o[pixel * 3 + 2] = ray.o.z;
  out_d[pixel * 3 + 0] = ray.d.x, out_d[pixel * 3 + 1] = ray.d.y, out_d[pixel * 3 + 2] = ray.d.z;
}

__global__ void k_scatter(DMaterial m, const float *dirs, const float *pos, const float *nrm, const int32_t *front,
                          const float *u4, size_t n, int32_t *scattered, float *out_o, float *out_d, float *att_out,
                          float *emitted) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ray sr{v3(0, 0, 0), v3(0, 0, 0)};
  V3 att = v3(0, 0, 0);
  const V3 e = mat_emitted(m);
  const float u[4] = {u4[i * 4], u4[i * 4 + 1], u4[i * 4 + 2], u4[i * 4 + 3]};
  const bool ok = mat_scatter(m, v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]), v3(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]),
                              v3(nrm[i * 3], nrm[i * 3 + 1], nrm[i * 3 + 2]), front[i] != 0, u, sr, att);
  scattered[i] = ok ? 1 : 0;
  out_o[i * 3] = sr.o.x, out_o[i * 3 + 1] = sr.o.y, out_o[i * 3 + 2] = sr.o.z;
  out_d[i * 3] = sr.d.x, out_d[i * 3 + 1] = sr.d.y, out_d[i * 3 + 2] = sr.d.z;
  att_out[i * 3] = att.x, att_out[i * 3 + 1] = att.y, att_out[i * 3 + 2] = att.z;
  emitted[i * 3] = e.x, emitted[i * 3 + 1] = e.y, emitted[i * 3 + 2] = e.z;
}

__global__ void k_philox(U4 c, uint32_t k0, uint32_t k1, uint32_t *out) {
  const U4 r = philox4x32_10(c, k0, k1);
  out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    CK(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void upload(const T *src, size_t count) {
    alloc(count);
    if (count) CK(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};

}  // namespace

// ------------------------------------------------------------------------------------------------------------
struct ptc_scene {
  HostScene hs;
  bool committed = false;
  int device = -1;
  int sm_count = 0;
  DevBuf<DObject> d_objects;
  DevBuf<DMaterial> d_materials;
  DevBuf<DMesh> d_meshes;
  DevBuf<float> d_sky;
  std::vector<std::unique_ptr<DevBuf<float4>>> mesh_bufs;
  DScene ds;

  // render workspace
  uint32_t pool = 0;
  DevBuf<float4> w_ray_o[2], w_ray_d[2], w_beta[2], w_hit0, w_hit1;
  DevBuf<uint2> w_tq_ray[2];
  DevBuf<float4> w_tq_o[2], w_tq_d[2];
  DevBuf<float2> w_tq_res[2];
  DevBuf<uint32_t> w_round;  // RoundCtl storage: n[rounds + 1], cursor[rounds + 1]
  DevBuf<float> w_film, w_film_out;  // ptc_render / ptc_resolve_u32 staging, grow-only
  DevBuf<uint32_t> w_packed;
  int mesh_objects = 0;  // mesh entries in object_list = extend rounds needed
  DevBuf<Ctl> d_ctl;
  Ctl *h_ctl = nullptr;  // pinned ring
  static constexpr int kRing = 4;
  cudaEvent_t ring_ev[kRing] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t own_stream = nullptr;
  std::vector<cudaEvent_t> timing_events;

  ~ptc_scene() {
    if (device >= 0) cudaSetDevice(device);
    if (h_ctl) cudaFreeHost(h_ctl);
    for (auto &e : ring_ev)
      if (e) cudaEventDestroy(e);
    for (auto &e : timing_events) cudaEventDestroy(e);
    if (own_stream) cudaStreamDestroy(own_stream);
  }
};

namespace {

// One extend pass = pre, then (traverse, post) once per mesh object a ray can meet.  Returns the number of launches.
enum Stage { ST_PRE = 0, ST_TRAVERSE, ST_POST, ST_SHADE, ST_REGEN, ST_COUNT };

int launch_extend(cudaStream_t stream, int sm_count, Ctl *ctl, const RoundCtl &rc, const DScene &ds, const ExtendOut &eo,
                  const TaskQ &tq, int src, float t_min, float t_max, bool counters,
                  const std::function<void(int)> *mark = nullptr) {
  static const uint32_t refill = getenv("PTC_REFILL") ? (uint32_t)atoi(getenv("PTC_REFILL")) : kRefillLanes;   // tuning knobs
  static const int tblocks = getenv("PTC_TBLOCKS") ? atoi(getenv("PTC_TBLOCKS")) : 8;
  const dim3 grid(sm_count * 8), tgrid(sm_count * tblocks);
  int launches = 1;
  if (mark) (*mark)(ST_PRE);
  if (counters) k_extend_pre<true><<<grid, kExtendThreads, 0, stream>>>(ctl, rc, ds, eo, tq, src, t_min, t_max);
  else k_extend_pre<false><<<grid, kExtendThreads, 0, stream>>>(ctl, rc, ds, eo, tq, src, t_min, t_max);
  for (int r = 0; r < rc.rounds; r++) {
    if (mark) (*mark)(ST_TRAVERSE);
    if (counters) k_traverse<true><<<tgrid, kExtendThreads, 0, stream>>>(ctl, rc, ds, tq, r, t_min, refill);
    else k_traverse<false><<<tgrid, kExtendThreads, 0, stream>>>(ctl, rc, ds, tq, r, t_min, refill);
    if (mark) (*mark)(ST_POST);
    if (counters) k_extend_post<true><<<grid, kExtendThreads, 0, stream>>>(ctl, rc, ds, eo, tq, src, r, t_min, t_max);
    else k_extend_post<false><<<grid, kExtendThreads, 0, stream>>>(ctl, rc, ds, eo, tq, src, r, t_min, t_max);
    launches += 2;
  }
  return launches;
}

TaskQ taskq_of(ptc_scene *s);

// grow-only: `pool` slots of path state (128 B each) + task queues (96 B each, only for scenes with meshes)
void ensure_workspace(ptc_scene *s, uint32_t pool) {
  if (s->pool >= pool && s->h_ctl) return;
  for (int k = 0; k < 2; k++) {
    s->w_ray_o[k].alloc(pool);
    s->w_ray_d[k].alloc(pool);
    s->w_beta[k].alloc(pool);
  }
  s->w_hit0.alloc(pool);
  s->w_hit1.alloc(pool);
  if (s->mesh_objects > 0)
    for (int k = 0; k < 2; k++) {
      s->w_tq_ray[k].alloc(pool);
      s->w_tq_o[k].alloc(pool);
      s->w_tq_d[k].alloc(pool);
      s->w_tq_res[k].alloc(pool);
    }
  s->d_ctl.alloc(1);
  s->w_round.alloc(2 * (size_t)(s->mesh_objects + 1));
  if (!s->h_ctl) {
    CK(cudaMallocHost(&s->h_ctl, sizeof(Ctl) * ptc_scene::kRing));
    for (auto &e : s->ring_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  s->pool = pool;
}

Buffers buffers_of(ptc_scene *s) {
  Buffers b;
  for (int k = 0; k < 2; k++) {
    b.ray_o[k] = s->w_ray_o[k].p;
    b.ray_d[k] = s->w_ray_d[k].p;
    b.beta[k] = s->w_beta[k].p;
  }
  b.hit0 = s->w_hit0.p;
  b.hit1 = s->w_hit1.p;
  return b;
}

TaskQ taskq_of(ptc_scene *s) {
  TaskQ q;
  for (int k = 0; k < 2; k++) {
    q.ray[k] = s->w_tq_ray[k].p;
    q.o[k] = s->w_tq_o[k].p;
    q.d[k] = s->w_tq_d[k].p;
    q.res[k] = s->w_tq_res[k].p;
  }
  return q;
}

void require_committed(const ptc_scene *s) {
  if (!s) throw std::invalid_argument("null scene");
  if (!s->committed) throw std::logic_error("scene not committed (call ptc_scene_commit first)");
}

// the wavefront loop; adds radiance sums into d_accum on `stream`
void render_accumulate(ptc_scene *s, const ptc_camera *cam, const ptc_render_settings *st, float *d_accum,
                       cudaStream_t stream, ptc_stats *stats) {
  require_committed(s);
  if (!cam || !st || !d_accum) throw std::invalid_argument("null argument");
  if (st->width <= 0 || st->height <= 0 || st->spp <= 0) throw std::invalid_argument("bad render settings");
  if ((int64_t)st->width * st->height > (int64_t)0x3fffffff) throw std::invalid_argument("image too large");
  CK(cudaSetDevice(s->device));
  int s_begin = st->sample_begin, s_end = st->sample_end;
  if (s_begin == 0 && s_end == 0) s_end = st->spp;
  if (s_begin < 0 || s_end < s_begin) throw std::invalid_argument("bad sample range");
  const int tile_mod = st->tile_mod > 0 ? st->tile_mod : 1;
  const int tile_rem = st->tile_mod > 0 ? st->tile_rem : 0;
  if (tile_rem < 0 || tile_rem >= tile_mod) throw std::invalid_argument("tile_rem out of range");
  const uint64_t want_paths = (uint64_t)st->width * st->height * (uint64_t)(s_end - s_begin);
  uint32_t pool;
  if (st->pool_paths > 0) {
    pool = std::max((uint32_t)st->pool_paths, 1024u);
  } else {  // default: 4 M slots (measured best on B200 for config C2), less for renders that cannot fill them
    pool = 1u << 16;
    while (pool < (1u << 22) && pool < want_paths) pool <<= 1;
  }
  ensure_workspace(s, pool);

  RenderParams rp;
  memcpy(&rp.cam, cam, sizeof(DCamera));
  rp.width = st->width, rp.height = st->height;
  rp.max_depth = st->max_depth;
  rp.sample_begin = s_begin, rp.n_samples = s_end - s_begin;
  rp.tiles_x = (st->width + 31) / 32;
  const int tiles_y = (st->height + 31) / 32;
  const int n_tiles = rp.tiles_x * tiles_y;
  rp.n_my_tiles = tile_rem < n_tiles ? (n_tiles - tile_rem + tile_mod - 1) / tile_mod : 0;
  rp.tile_mod = tile_mod, rp.tile_rem = tile_rem;
  rp.pool = pool;
  rp.seed = st->seed;

  // valid pixels of this rank (for the paths statistic)
  uint64_t my_pixels = 0;
  for (int t = tile_rem; t < n_tiles; t += tile_mod) {
    const int tx = t % rp.tiles_x, ty = t / rp.tiles_x;
    const int w = std::min(32, st->width - tx * 32), h = std::min(32, st->height - ty * 32);
    my_pixels += (uint64_t)w * h;
  }

  Ctl init;
  memset(&init, 0, sizeof(init));
  init.total_paths = rp.max_depth > 0 ? (unsigned long long)rp.n_my_tiles * 1024ull * (unsigned long long)rp.n_samples : 0ull;
  CK(cudaMemcpyAsync(s->d_ctl.p, &init, sizeof(Ctl), cudaMemcpyHostToDevice, stream));

  const bool counters = (st->flags & PTC_FLAG_COUNTERS) != 0;
  const bool timing = (st->flags & PTC_FLAG_TIMING) != 0;
  const Buffers b = buffers_of(s);
  const TaskQ tq = taskq_of(s);
  const RoundCtl rc{s->w_round.p, s->w_round.p + s->mesh_objects + 1, s->mesh_objects};
  const int sms = s->sm_count;
  const dim3 g_shade(sms * 3), g_gen(sms * 4);

  cudaEvent_t ev_begin, ev_end;
  CK(cudaEventCreate(&ev_begin));
  CK(cudaEventCreate(&ev_end));
  size_t tev_used = 0;
  auto tev = [&]() -> cudaEvent_t {
    if (tev_used == s->timing_events.size()) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      s->timing_events.push_back(e);
    }
    return s->timing_events[tev_used++];
  };

  std::vector<int> mark_stage;
  const std::function<void(int)> mark = [&](int stage) {
    CK(cudaEventRecord(tev(), stream));
    mark_stage.push_back(stage);
  };
  CK(cudaEventRecord(ev_begin, stream));
  int cur = 0;
  k_advance<<<1, 32, 0, stream>>>(s->d_ctl.p, rc, pool);
  k_generate<<<g_gen, 256, 0, stream>>>(s->d_ctl.p, rp, b, cur);
  uint64_t launches = 2;
  std::deque<int> pending;
  int ring_next = 0;
  uint64_t it = 0;
  const int check_every = 4;
  bool finished = init.total_paths == 0;
  while (!finished) {
    const ExtendOut eo{b, nullptr};
    launches += launch_extend(stream, sms, s->d_ctl.p, rc, s->ds, eo, tq, cur, kEps, INFINITY, counters, timing ? &mark : nullptr) - 1;
    if (timing) mark(ST_SHADE);
    k_shade<<<g_shade, kShadeThreads, kShadeSmem, stream>>>(s->d_ctl.p, s->ds, rp, b, cur, cur ^ 1, d_accum);
    if (timing) mark(ST_REGEN);
    k_advance<<<1, 32, 0, stream>>>(s->d_ctl.p, rc, pool);
    k_generate<<<g_gen, 256, 0, stream>>>(s->d_ctl.p, rp, b, cur ^ 1);
    launches += 4;
    cur ^= 1;
    it++;
    if (it % check_every == 0) {
      // snapshot the control block; consume finished snapshots without stalling the launch queue.  The host only
      // blocks when the ring is full, i.e. when it is kRing * check_every iterations ahead of what it has seen.
      const int k = ring_next;
      ring_next = (ring_next + 1) % ptc_scene::kRing;
      CK(cudaMemcpyAsync(&s->h_ctl[k], s->d_ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
      CK(cudaEventRecord(s->ring_ev[k], stream));
      pending.push_back(k);
      while (!pending.empty()) {
        const int o = pending.front();
        const bool must_wait = (int)pending.size() >= ptc_scene::kRing;
        if (!must_wait && cudaEventQuery(s->ring_ev[o]) == cudaErrorNotReady) {
          cudaGetLastError();
          break;
        }
        CK(cudaEventSynchronize(s->ring_ev[o]));
        pending.pop_front();
        if (s->h_ctl[o].done) {
          finished = true;
          break;
        }
      }
    }
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(ev_end, stream));
  Ctl fin;
  CK(cudaMemcpyAsync(&s->h_ctl[0], s->d_ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  fin = s->h_ctl[0];
  if (!fin.done) throw std::runtime_error("wavefront loop ended before all paths terminated");
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->paths = rp.max_depth > 0 ? my_pixels * (uint64_t)rp.n_samples : 0;
    stats->rays = fin.rays;
    stats->iterations = fin.iterations;
    stats->kernel_launches = launches;
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
    stats->render_ms = ms;
    if (timing) {
      // every mark opens a stage that lasts until the next mark (the last one until the end-of-render event)
      double per[ST_COUNT] = {0, 0, 0, 0, 0};
      uint64_t n_ext = 0;
      for (size_t k = 0; k < tev_used; k++) {
        float a = 0.0f;
        CK(cudaEventElapsedTime(&a, s->timing_events[k], k + 1 < tev_used ? s->timing_events[k + 1] : ev_end));
        per[mark_stage[k]] += a;
        if (mark_stage[k] == ST_PRE) n_ext++;
      }
      stats->pre_ms = per[ST_PRE];
      stats->traverse_ms = per[ST_TRAVERSE];
      stats->post_ms = per[ST_POST];
      stats->extend_ms = per[ST_PRE] + per[ST_TRAVERSE] + per[ST_POST];
      stats->shade_ms = per[ST_SHADE];
      stats->regen_ms = per[ST_REGEN];
      stats->extend_launches = n_ext;
    }
    stats->nodes_visited = fin.nodes;
    stats->tris_tested = fin.tris;
    stats->mesh_rays = fin.mesh_rays;
  }
  cudaEventDestroy(ev_begin);
  cudaEventDestroy(ev_end);
}

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define PTC_GUARD_BEGIN try {
#define PTC_GUARD_END                                   \
  }                                                     \
  catch (CudaError & e) { return fail(PTC_E_CUDA, e.what()); }             \
  catch (std::bad_alloc & e) { return fail(PTC_E_NOMEM, e.what()); }       \
  catch (std::logic_error & e) {                                           \
    return fail(dynamic_cast<std::invalid_argument *>(&e) ? PTC_E_INVALID : PTC_E_STATE, e.what()); \
  }                                                                        \
  catch (std::exception & e) { return fail(PTC_E_INVALID, e.what()); }

}  // namespace

// =============================================================================================================
extern "C" {

const char *ptc_last_error(void) { return g_err.c_str(); }
int ptc_abi_version(void) { return PTC_ABI_VERSION; }
int ptc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

ptc_scene *ptc_scene_create(void) {
  try {
    return new ptc_scene();
  } catch (std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void ptc_scene_destroy(ptc_scene *s) { delete s; }

int ptc_scene_add_material(ptc_scene *s, const ptc_material *m) {
  PTC_GUARD_BEGIN
  if (!s || !m) throw std::invalid_argument("null argument");
  if (s->committed) throw std::logic_error("scene already committed");
  return s->hs.add_material(m);
  PTC_GUARD_END
}
int ptc_scene_add_sphere(ptc_scene *s, const float c[3], float radius, int material) {
  PTC_GUARD_BEGIN
  if (!s || !c) throw std::invalid_argument("null argument");
  if (s->committed) throw std::logic_error("scene already committed");
  return s->hs.add_sphere(c, radius, material);
  PTC_GUARD_END
}
int ptc_scene_add_plane(ptc_scene *s, const float p1[3], const float n[3], int material) {
  PTC_GUARD_BEGIN
  if (!s || !p1 || !n) throw std::invalid_argument("null argument");
  if (s->committed) throw std::logic_error("scene already committed");
  return s->hs.add_plane(p1, n, material);
  PTC_GUARD_END
}
int ptc_scene_add_quad(ptc_scene *s, const float base[3], const float e0[3], const float e1[3], const float n[3], float d,
                       float inv0, float inv1, int material) {
  PTC_GUARD_BEGIN
  if (!s || !base || !e0 || !e1 || !n) throw std::invalid_argument("null argument");
  if (s->committed) throw std::logic_error("scene already committed");
  return s->hs.add_quad(base, e0, e1, n, d, inv0, inv1, material);
  PTC_GUARD_END
}
int ptc_scene_add_cube(ptc_scene *s, const float o2w[16], const float w2o[16], int material) {
  PTC_GUARD_BEGIN
  if (!s || !o2w || !w2o) throw std::invalid_argument("null argument");
  if (s->committed) throw std::logic_error("scene already committed");
  return s->hs.add_cube(o2w, w2o, material);
  PTC_GUARD_END
}
int ptc_scene_add_mesh(ptc_scene *s, const float *tris, int64_t n, const float o2w[16], const float w2o[16], int material) {
  PTC_GUARD_BEGIN
  if (!s || !o2w || !w2o) throw std::invalid_argument("null argument");
  if (s->committed) throw std::logic_error("scene already committed");
  return s->hs.add_mesh(tris, n, o2w, w2o, material);
  PTC_GUARD_END
}
int ptc_scene_set_sky_hdr(ptc_scene *s, const float *rgb, int32_t w, int32_t h) {
  PTC_GUARD_BEGIN
  if (!s) throw std::invalid_argument("null argument");
  if (s->committed) throw std::logic_error("scene already committed");
  s->hs.set_sky(rgb, w, h);
  return 0;
  PTC_GUARD_END
}

int ptc_scene_build(ptc_scene *s) {
  PTC_GUARD_BEGIN
  if (!s) throw std::invalid_argument("null argument");
  s->hs.build_all();
  return 0;
  PTC_GUARD_END
}

int ptc_scene_commit(ptc_scene *s, int device) {
  PTC_GUARD_BEGIN
  if (!s) throw std::invalid_argument("null argument");
  if (s->committed) throw std::logic_error("scene already committed");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    throw CudaError("no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= n) throw std::invalid_argument("device index out of range");
  s->hs.build_all();  // host flattening: reference-BVH dead mask -> SAH -> 8-wide quantised BVH
  CK(cudaSetDevice(device));
  s->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  s->sm_count = prop.multiProcessorCount;
  std::vector<DMesh> dm;
  for (auto &m : s->hs.meshes) {
    auto nodes = std::make_unique<DevBuf<float4>>();
    auto tris = std::make_unique<DevBuf<float4>>();
    auto nrm = std::make_unique<DevBuf<float4>>();
    nodes->upload(reinterpret_cast<const float4 *>(m->nodes.data()), m->nodes.size() * 5);
    tris->upload(reinterpret_cast<const float4 *>(m->tri48.data()), m->tri48.size() * 3);
    nrm->upload(m->normals.data(), m->normals.size());
    const DMesh d = make_dmesh(*m, nodes->p, tris->p, nrm->p);
    dm.push_back(d);
    s->mesh_bufs.push_back(std::move(nodes));
    s->mesh_bufs.push_back(std::move(tris));
    s->mesh_bufs.push_back(std::move(nrm));
  }
  s->d_objects.upload(s->hs.objects.data(), s->hs.objects.size());
  s->d_materials.upload(s->hs.materials.data(), s->hs.materials.size());
  s->d_meshes.upload(dm.data(), dm.size());
  if (!s->hs.sky.empty()) s->d_sky.upload(s->hs.sky.data(), s->hs.sky.size());
  s->ds.objects = s->d_objects.p;
  s->ds.materials = s->d_materials.p;
  s->ds.meshes = s->d_meshes.p;
  s->ds.sky = s->hs.sky.empty() ? nullptr : s->d_sky.p;
  s->ds.n_objects = (int32_t)s->hs.objects.size();
  s->ds.n_materials = (int32_t)s->hs.materials.size();
  s->ds.n_meshes = (int32_t)dm.size();
  s->ds.sky_w = s->hs.sky_w;
  s->ds.sky_h = s->hs.sky_h;
  s->mesh_objects = 0;
  for (const DObject &o : s->hs.objects)
    if (o.type == OBJ_MESH) s->mesh_objects++;
  CK(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
  s->committed = true;
  return 0;
  PTC_GUARD_END
}

int ptc_scene_mesh_info(const ptc_scene *s, int object, ptc_mesh_info *info, uint8_t *dead, int32_t *order) {
  PTC_GUARD_BEGIN
  if (!s) throw std::invalid_argument("null argument");
  if (object < 0 || (size_t)object >= s->hs.objects.size() || s->hs.objects[(size_t)object].type != OBJ_MESH)
    throw std::invalid_argument("object is not a mesh");
  const MeshBuild &m = *s->hs.meshes[(size_t)s->hs.objects[(size_t)object].mesh];
  if (!m.built) throw std::logic_error("mesh not built yet (call ptc_scene_build or ptc_scene_commit first)");
  if (info) {
    info->triangles = m.n;
    info->live_triangles = m.live;
    info->ref_nodes = m.ref_nodes;
    info->ref_leaves = m.ref_leaves;
    info->ref_depth = m.ref_depth;
    info->wide_nodes = (int64_t)m.nodes.size();
    info->wide_depth = m.wide_depth;
    info->node_bytes = (int64_t)m.nodes.size() * (int64_t)sizeof(Node8);
    info->triangle_bytes = (int64_t)m.tri48.size() * (int64_t)sizeof(Tri48);
  }
  if (dead) memcpy(dead, m.dead.data(), (size_t)m.n);
  if (order) memcpy(order, m.order.data(), (size_t)m.n * sizeof(int32_t));
  return 0;
  PTC_GUARD_END
}

int ptc_render_accumulate(ptc_scene *s, const ptc_camera *cam, const ptc_render_settings *st, float *d_accum,
                          void *cuda_stream, ptc_stats *stats) {
  PTC_GUARD_BEGIN
  require_committed(s);
  render_accumulate(s, cam, st, d_accum, cuda_stream ? (cudaStream_t)cuda_stream : s->own_stream, stats);
  return 0;
  PTC_GUARD_END
}

int ptc_render(ptc_scene *s, const ptc_camera *cam, const ptc_render_settings *st, float *out_rgb, ptc_stats *stats) {
  PTC_GUARD_BEGIN
  require_committed(s);
  if (!st || !out_rgb) throw std::invalid_argument("null argument");
  if (st->width <= 0 || st->height <= 0 || st->spp <= 0) throw std::invalid_argument("bad render settings");
  CK(cudaSetDevice(s->device));
  const size_t n = (size_t)st->width * st->height * 3;
  if (s->w_film.n < n) s->w_film.alloc(n);
  if (s->w_film_out.n < n) s->w_film_out.alloc(n);
  DevBuf<float> &accum = s->w_film, &out = s->w_film_out;
  cudaStream_t stream = s->own_stream;
  CK(cudaMemsetAsync(accum.p, 0, n * sizeof(float), stream));
  render_accumulate(s, cam, st, accum.p, stream, stats);
  const float inv_spp = 1.0f / (float)st->spp;  // renderer.rs:85
  k_scale<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(accum.p, out.p, n, inv_spp);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_rgb, out.p, n * sizeof(float), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  if (stats) stats->kernel_launches += 1;
  return 0;
  PTC_GUARD_END
}

int ptc_resolve_device(const float *d_rgb, int64_t n_pixels, float scale, uint32_t *d_out, void *cuda_stream) {
  PTC_GUARD_BEGIN
  if (!d_rgb || !d_out || n_pixels < 0) throw std::invalid_argument("bad argument");
  if (n_pixels == 0) return 0;
  k_resolve<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(d_rgb, (size_t)n_pixels, scale, d_out);
  CK(cudaGetLastError());
  return 0;
  PTC_GUARD_END
}

int ptc_resolve_u32(ptc_scene *s, const float *rgb, int64_t n_pixels, float scale, uint32_t *out) {
  PTC_GUARD_BEGIN
  require_committed(s);
  if (!rgb || !out || n_pixels < 0) throw std::invalid_argument("bad argument");
  if (n_pixels == 0) return 0;
  CK(cudaSetDevice(s->device));
  if (s->w_film_out.n < (size_t)n_pixels * 3) s->w_film_out.alloc((size_t)n_pixels * 3);
  if (s->w_packed.n < (size_t)n_pixels) s->w_packed.alloc((size_t)n_pixels);
  DevBuf<float> &d_in = s->w_film_out;
  DevBuf<uint32_t> &d_out = s->w_packed;
  CK(cudaMemcpyAsync(d_in.p, rgb, (size_t)n_pixels * 3 * sizeof(float), cudaMemcpyHostToDevice, s->own_stream));
  k_resolve<<<(unsigned)((n_pixels + 255) / 256), 256, 0, s->own_stream>>>(d_in.p, (size_t)n_pixels, scale, d_out.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d_out.p, (size_t)n_pixels * 4, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaStreamSynchronize(s->own_stream));
  return 0;
  PTC_GUARD_END
}

int ptc_intersect(ptc_scene *s, const float *origins, const float *dirs, int64_t n, float t_min, float t_max, ptc_hit *out,
                  ptc_stats *stats) {
  PTC_GUARD_BEGIN
  require_committed(s);
  if (n < 0 || (n > 0 && (!origins || !dirs || !out))) throw std::invalid_argument("bad argument");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n == 0) return 0;
  CK(cudaSetDevice(s->device));
  if (n > (int64_t)0x7fffffff) throw std::invalid_argument("too many rays for one call");
  DevBuf<float> d_o, d_d;
  DevBuf<ptc_hit> d_out;
  DevBuf<float4> ro, rd, h0, h1;
  DevBuf<int2> ids;
  DevBuf<Ctl> ctl;
  d_o.upload(origins, (size_t)n * 3);
  d_d.upload(dirs, (size_t)n * 3);
  d_out.alloc((size_t)n);
  ro.alloc((size_t)n), rd.alloc((size_t)n), h0.alloc((size_t)n), h1.alloc((size_t)n), ids.alloc((size_t)n);
  Ctl init;
  memset(&init, 0, sizeof(init));
  init.n_cur = (uint32_t)n;
  ctl.upload(&init, 1);
  cudaStream_t st = s->own_stream;
  const unsigned nb = (unsigned)((n + 255) / 256);
  k_pack_rays<<<nb, 256, 0, st>>>(d_o.p, d_d.p, (size_t)n, ro.p, rd.p);
  ExtendOut eo;
  eo.b.ray_o[0] = eo.b.ray_o[1] = ro.p;
  eo.b.ray_d[0] = eo.b.ray_d[1] = rd.p;
  eo.b.beta[0] = eo.b.beta[1] = nullptr;
  eo.b.hit0 = h0.p, eo.b.hit1 = h1.p;
  eo.ids = ids.p;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a, st));
  DevBuf<uint2> q_ray[2];
  DevBuf<float4> q_o[2], q_d[2];
  DevBuf<float2> q_res[2];
  TaskQ tq;
  for (int k = 0; k < 2; k++) {
    const size_t cap = s->mesh_objects > 0 ? (size_t)n : 1;
    q_ray[k].alloc(cap), q_o[k].alloc(cap), q_d[k].alloc(cap), q_res[k].alloc(cap);
    tq.ray[k] = q_ray[k].p, tq.o[k] = q_o[k].p, tq.d[k] = q_d[k].p, tq.res[k] = q_res[k].p;
  }
  DevBuf<uint32_t> round;
  round.alloc(2 * (size_t)(s->mesh_objects + 1));
  CK(cudaMemsetAsync(round.p, 0, round.n * sizeof(uint32_t), st));
  const RoundCtl rc{round.p, round.p + s->mesh_objects + 1, s->mesh_objects};
  const int ext_launches = launch_extend(st, s->sm_count, ctl.p, rc, s->ds, eo, tq, 0, t_min, t_max, true);
  CK(cudaGetLastError());
  CK(cudaEventRecord(b, st));
  k_unpack_hits<<<nb, 256, 0, st>>>(h0.p, h1.p, ids.p, (size_t)n, d_out.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d_out.p, (size_t)n * sizeof(ptc_hit), cudaMemcpyDeviceToHost, s->own_stream));
  Ctl fin;
  CK(cudaMemcpyAsync(&fin, ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaStreamSynchronize(s->own_stream));
  const unsigned long long ctr[3] = {fin.nodes, fin.tris, fin.mesh_rays};
  if (stats) {
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, a, b));
    stats->rays = (uint64_t)n;
    stats->kernel_launches = 2 + (uint64_t)ext_launches;
    stats->render_ms = ms;
    stats->extend_ms = ms;
    stats->extend_launches = 1;
    stats->nodes_visited = ctr[0];
    stats->tris_tested = ctr[1];
    stats->mesh_rays = ctr[2];
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return 0;
  PTC_GUARD_END
}

int ptc_primary_rays(ptc_scene *s, const ptc_camera *cam, const ptc_render_settings *st, int32_t sample, float *out_o,
                     float *out_d) {
  PTC_GUARD_BEGIN
  require_committed(s);
  if (!cam || !st || !out_o || !out_d || st->width <= 0 || st->height <= 0) throw std::invalid_argument("bad argument");
  CK(cudaSetDevice(s->device));
  RenderParams rp;
  memset(&rp, 0, sizeof(rp));
  memcpy(&rp.cam, cam, sizeof(DCamera));
  rp.width = st->width, rp.height = st->height, rp.seed = st->seed;
  const size_t n = (size_t)st->width * st->height;
  DevBuf<float> d_o, d_d;
  d_o.alloc(n * 3);
  d_d.alloc(n * 3);
  k_primary_rays<<<(unsigned)((n + 255) / 256), 256, 0, s->own_stream>>>(rp, (uint32_t)sample, d_o.p, d_d.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_o, d_o.p, n * 12, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaMemcpyAsync(out_d, d_d.p, n * 12, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaStreamSynchronize(s->own_stream));
  return 0;
  PTC_GUARD_END
}

int ptc_scatter(ptc_scene *s, int material, const float *ray_dirs, const float *positions, const float *normals,
                const int32_t *front_face, const float *u4, int64_t n, int32_t *scattered, float *out_origin, float *out_dir,
                float *attenuation, float *emitted) {
  PTC_GUARD_BEGIN
  require_committed(s);
  if (material < 0 || (size_t)material >= s->hs.materials.size()) throw std::invalid_argument("material index out of range");
  if (n < 0) throw std::invalid_argument("bad argument");
  if (n == 0) return 0;
  CK(cudaSetDevice(s->device));
  const size_t N = (size_t)n;
  DevBuf<float> d_dir, d_pos, d_nrm, d_u, d_oo, d_od, d_att, d_em;
  DevBuf<int32_t> d_front, d_sc;
  d_dir.upload(ray_dirs, N * 3);
  d_pos.upload(positions, N * 3);
  d_nrm.upload(normals, N * 3);
  d_u.upload(u4, N * 4);
  d_front.upload(front_face, N);
  d_oo.alloc(N * 3), d_od.alloc(N * 3), d_att.alloc(N * 3), d_em.alloc(N * 3), d_sc.alloc(N);
  k_scatter<<<(unsigned)((N + 127) / 128), 128, 0, s->own_stream>>>(s->hs.materials[(size_t)material], d_dir.p, d_pos.p, d_nrm.p,
                                                                    d_front.p, d_u.p, N, d_sc.p, d_oo.p, d_od.p, d_att.p, d_em.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(scattered, d_sc.p, N * 4, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaMemcpyAsync(out_origin, d_oo.p, N * 12, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaMemcpyAsync(out_dir, d_od.p, N * 12, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaMemcpyAsync(attenuation, d_att.p, N * 12, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaMemcpyAsync(emitted, d_em.p, N * 12, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaStreamSynchronize(s->own_stream));
  return 0;
  PTC_GUARD_END
}

int ptc_philox(ptc_scene *s, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  PTC_GUARD_BEGIN
  require_committed(s);
  if (!ctr || !key || !out) throw std::invalid_argument("null argument");
  CK(cudaSetDevice(s->device));
  DevBuf<uint32_t> d;
  d.alloc(4);
  k_philox<<<1, 1, 0, s->own_stream>>>(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1], d.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d.p, 16, cudaMemcpyDeviceToHost, s->own_stream));
  CK(cudaStreamSynchronize(s->own_stream));
  return 0;
  PTC_GUARD_END
}

}  // extern "C"
