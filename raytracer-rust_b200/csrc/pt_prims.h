// pt_prims.h — closest hit over the object list: HittableList::hit (src/hittable.rs:46-57) and the `hit` of every
// primitive, with the reference's interval conventions (open for sphere / plane / quad / triangle, closed world-t
// recheck for cube and mesh) and operation order.
#pragma once
#include "pt_bvh8.h"
#include "pt_math.h"
#include "pt_types.h"

namespace pt {

// hittable.rs:19-26
PT_HD void set_face_normal(Hit &h, V3 ray_d, V3 outward) {
  bool front = dot(ray_d, outward) < 0.0f;
  V3 n = front ? outward : -outward;
  h.front_face = front ? 1 : 0;
  h.nx = n.x;
  h.ny = n.y;
  h.nz = n.z;
}
PT_HD void set_pos(Hit &h, V3 p) {
  h.px = p.x;
  h.py = p.y;
  h.pz = p.z;
}

// src/objects/sphere.rs:15-53
PT_HD bool hit_sphere(const float *f, const Ray &ray, float t_min, float t_max, Hit &h) {
  const float4 cr = ld4(reinterpret_cast<const float4 *>(f));
  const V3 center = v3(cr.x, cr.y, cr.z);
  const float radius = cr.w;
  const V3 oc = ray.o - center;
  const float a = dot(ray.d, ray.d);
  const float half_b = dot(oc, ray.d);
  const float c = dot(oc, oc) - radius * radius;
  const float disc = half_b * half_b - a * c;
  if (disc < 0.0f) return false;
  const float sqrtd = sqrtf(disc);
  float root = (-half_b - sqrtd) / a;
  if (root <= t_min || root >= t_max) {
    root = (-half_b + sqrtd) / a;
    if (root <= t_min || root >= t_max) return false;
  }
  h.t = root;
  const V3 p = ray_at(ray, root);
  set_pos(h, p);
  const V3 outward = (p - center) / radius;  // vec3.rs:117-122 panics for |radius| < 1e-4: such spheres are refused by ptc_scene_add_sphere
  set_face_normal(h, ray.d, outward);
  return true;
}

// src/objects/plane.rs:26-56
PT_HD bool hit_plane(const float *f, const Ray &ray, float t_min, float t_max, Hit &h) {
  const V3 p1 = v3(f[0], f[1], f[2]), n = v3(f[3], f[4], f[5]);
  const float denom = dot(n, ray.d);
  if (fabsf(denom) < kEps) return false;
  const float t = dot(n, p1 - ray.o) / denom;
  if (t <= t_min || t >= t_max) return false;
  h.t = t;
  set_pos(h, ray_at(ray, t));
  set_face_normal(h, ray.d, n);
  return true;
}

// src/tungsten/objects/quad.rs:83-132
PT_HD bool hit_quad(const float *f, const Ray &ray, float t_min, float t_max, Hit &h) {
  const float4 *f4 = reinterpret_cast<const float4 *>(f);
  const float4 q0 = ld4(f4), q1 = ld4(f4 + 1), q2 = ld4(f4 + 2), q3 = ld4(f4 + 3);
  const V3 base = v3(q0.x, q0.y, q0.z), e0 = v3(q0.w, q1.x, q1.y), e1 = v3(q1.z, q1.w, q2.x), n = v3(q2.y, q2.z, q2.w);
  const float d = q3.x, inv0 = q3.y, inv1 = q3.z;
  const float denom = dot(n, ray.d);
  if (fabsf(denom) < kEps) return false;
  const float t = (d - dot(n, ray.o)) / denom;
  if (t <= t_min || t >= t_max) return false;
  const V3 p = ray_at(ray, t);
  const V3 v = p - base;
  const float l0 = dot(v, e0) * inv0;
  const float l1 = dot(v, e1) * inv1;
  const float lo = -kEps, hi = 1.0f + kEps;
  if (!((lo <= l0 && l0 <= hi) && (lo <= l1 && l1 <= hi))) return false;
  h.t = t;
  set_pos(h, p);
  set_face_normal(h, ray.d, n);
  return true;
}

// src/objects/cube.rs:59-158.  f[0..15] = world_to_object, f[16..31] = object_to_world.
//
// Cube::hit in two halves.  Whether a cube hit is accepted depends on t_obj and t_world only; the normal (face pick,
// normalize_or_zero, (M^-1)^T n, normalize, face flip: more than half of the instructions of an accepted hit) is a pure
// function of the object-space hit point and is only ever read from the record that wins the list scan.  hit_cube_t
// does the first half and leaves the object-space point in the record's normal fields with front_face =
// kCubeNormalPending; cube_finish_normal completes the record.  The extend stage calls the second half once per ray,
// after the scan, instead of once per accepted cube inside it (in C2 a warp of 32 incoherent rays used to run the full
// hit path of all three cubes at ~10 active lanes each).  hit_cube = both halves back to back, same bits.
constexpr int32_t kCubeNormalPending = 2;
PT_HD bool hit_cube_t(const float *f, const Ray &ray, float t_min, float t_max, Hit &h) {
  const M4 w2o = load_m4(f);
  V3 o, d;
  mat_point_vector(w2o, ray.o, ray.d, o, d);  // d not renormalised (cube.rs:70-83)
  const float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
#if defined(__CUDA_ARCH__)
  // (-0.5 - o) * i and (0.5 - o) * i, x and y as packed pairs (pt_math.h: same roundings as the scalar statements)
  const F2 a1 = sub2(-0.5f, -0.5f, o.x, o.y), a2 = sub2(0.5f, 0.5f, o.x, o.y);
  const F2 m1 = mul2(a1.x, a1.y, ix, iy), m2 = mul2(a2.x, a2.y, ix, iy);
  const float t1x = m1.x, t1y = m1.y, t2x = m2.x, t2y = m2.y;
#else
  const float t1x = (-0.5f - o.x) * ix, t2x = (0.5f - o.x) * ix;
  const float t1y = (-0.5f - o.y) * iy, t2y = (0.5f - o.y) * iy;
#endif
  const float t1z = (-0.5f - o.z) * iz, t2z = (0.5f - o.z) * iz;
  const float t_enter = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
  const float t_exit = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
  if (t_exit < t_enter || t_exit <= 0.0f) return false;
  const float t_obj = t_enter > 0.0f ? t_enter : t_exit;
  if (t_obj >= t_max || t_obj <= t_min || t_obj < kEps) return false;
  const V3 p = o + d * t_obj;
  const M4 o2w = load_m4(f + 16);
  const V3 pw = mat_point_packed(o2w, p);
  if (dot(pw - ray.o, ray.d) < 0.0f) return false;
  const float t_world = dot(pw - ray.o, ray.d);
  if (t_world < t_min || t_world > t_max) return false;  // closed interval (cube.rs:150)
  h.t = t_world;
  set_pos(h, pw);
  h.nx = p.x, h.ny = p.y, h.nz = p.z;  // object-space hit point, for cube_finish_normal
  h.front_face = kCubeNormalPending;
  return true;
}
PT_HD void cube_finish_normal(const float *f, const Ray &ray, Hit &h) {
  const V3 p = v3(h.nx, h.ny, h.nz);
  V3 n = v3(0, 0, 0);
  const float ax = fabsf(p.x), ay = fabsf(p.y), az = fabsf(p.z);
  const float tol = 1e-4f;
  if (fabsf(ax - 0.5f) < tol) n.x = signum(p.x);
  else if (fabsf(ay - 0.5f) < tol) n.y = signum(p.y);
  else if (fabsf(az - 0.5f) < tol) n.z = signum(p.z);
  else if (ax > ay && ax > az) n.x = signum(p.x);
  else if (ay > az) n.y = signum(p.y);
  else n.z = signum(p.z);
  {  // glam normalize_or_zero.  n is +-1 on one axis unless p held a NaN: |n|^2 == 1 exactly, 1/sqrt(1) == 1, n * 1 == n
    const float nn = dot(n, n);
    if (nn != 1.0f) {
      const float rcp = 1.0f / sqrtf(nn);
      if (!isinf_f(rcp) && !isnan_f(rcp) && rcp > 0.0f) n = n * rcp;
      else n = v3(0, 0, 0);
    }
  }
  const M4 w2o = load_m4(f);
  const V3 nw = normalized(mat_t_vector(w2o, n));
  set_face_normal(h, ray.d, nw);
}
PT_HD bool hit_cube(const float *f, const Ray &ray, float t_min, float t_max, Hit &h) {
  if (!hit_cube_t(f, ray, t_min, t_max, h)) return false;
  cube_finish_normal(f, ray, h);
  return true;
}

// Object-space ray of Mesh::hit (mesh_object.rs:264-287): d_raw = M^-1 (d,0); d = that, normalised at :287 and again by
// Ray::new (ray.rs:15).
struct MeshRay {
  V3 o, d_raw, d;
};
template <bool PACKED = true>
PT_HD MeshRay mesh_object_ray(const float *w2o_f, const Ray &ray) {
  const M4 w2o = load_m4(w2o_f);
  MeshRay r;
  mat_point_vector<PACKED>(w2o, ray.o, ray.d, r.o, r.d_raw);
  r.d = normalized(normalized(r.d_raw));
  return r;
}
// Conservative pre-test against the padded frame of the root node: false only if no live triangle can be hit in
// (t_min, t_max).  Same NaN-dropping slab arithmetic as the node test in pt_bvh8.h.
PT_HD bool mesh_root_may_hit(const DMesh &mesh, const MeshRay &r, float t_min, float t_max) {
  const float idx = box_rcp(r.d.x), idy = box_rcp(r.d.y), idz = box_rcp(r.d.z);  // conservative test: no IEEE division needed
  const float ax = (mesh.root_lo[0] - r.o.x) * idx, bx = (mesh.root_hi[0] - r.o.x) * idx;
  const float ay = (mesh.root_lo[1] - r.o.y) * idy, by = (mesh.root_hi[1] - r.o.y) * idy;
  const float az = (mesh.root_lo[2] - r.o.z) * idz, bz = (mesh.root_hi[2] - r.o.z) * idz;
  const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t_min));
  const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t_max));
  return tn <= tf;
}

// Second half of Mesh::hit (mesh_object.rs:293-326): object-space closest triangle -> world-space HitRecord, with the
// closed-interval recheck on the (quirky) world t.
// `nq` = mesh.normals[mh.order], loaded by the caller (k_extend_post issues it together with its other gathers).
PT_HD bool mesh_finish(const float *f, float4 nq, const Ray &ray, const MeshRay &mr, const MeshHit &mh, float t_min, float t_max,
                       Hit &h) {
  const M4 w2o = load_m4(f), o2w = load_m4(f + 16);
  const V3 o = mr.o, d = mr.d;
  const V3 p_obj = o + d * mh.t;  // ray.at(t), bvh.rs:119
  V3 n_obj = v3(nq.x, nq.y, nq.z);
  if (!(dot(d, n_obj) < 0.0f)) n_obj = -n_obj;  // bvh.rs:120-126
  const V3 pw = mat_point(o2w, p_obj);
  const V3 nw = normalized(mat_t_vector(w2o, n_obj));
  const float t_world = mh.t * length(mr.d_raw) / length(ray.d);  // mesh_object.rs:312-314
  if (t_world < t_min || t_world > t_max) return false;             // closed (mesh_object.rs:316)
  h.t = t_world;
  h.triangle = (int32_t)mh.tri;
  set_pos(h, pw);
  set_face_normal(h, ray.d, nw);
  return true;
}

PT_HD bool mesh_finish(const float *f, const DMesh &mesh, const Ray &ray, const MeshRay &mr, const MeshHit &mh, float t_min,
                       float t_max, Hit &h) {
  return mesh_finish(f, ldg4(mesh.normals + mh.order), ray, mr, mh, t_min, t_max, h);
}

// src/mesh/mesh_object.rs:262-329, straight line (parity hooks and the hostsim harness; the renderer splits it into
// root test / traversal / finish across kernels, ptcore.cu)
template <bool COUNT>
PT_HD bool hit_mesh(const float *f, const DMesh &mesh, const Ray &ray, float t_min, float t_max, Hit &h,
                    TraversalCounters *ctr) {
  const MeshRay mr = mesh_object_ray(f, ray);
  if (!mesh_root_may_hit(mesh, mr, t_min, t_max)) return false;
  if (COUNT) ctr->mesh_rays++;
  MeshHit mh;
  // the BVH is queried with the WORLD t bounds (mesh_object.rs:289-291)
  if (!bvh8_closest<COUNT>(mesh, mr.o, mr.d, t_min, t_max, mh, ctr)) return false;
  return mesh_finish(f, mesh, ray, mr, mh, t_min, t_max, h);
}

// HittableList::hit (hittable.rs:46-57): linear scan in insertion order with a shrinking t_max; each primitive's
// own interval test decides what happens on equal t.
template <bool COUNT>
PT_HD bool scene_hit(const DScene &sc, const Ray &ray, float t_min, float t_max, Hit &best, TraversalCounters *ctr) {
  float closest = t_max;
  bool any = false;
  best.object = -1;
  best.triangle = -1;
  best.material = -1;
  for (int i = 0; i < sc.n_objects; i++) {
    const DObject *ob = sc.objects + i;
    const int type = ob->type;
    Hit tmp;
    tmp.triangle = -1;
    bool hit;
    if (type == OBJ_SPHERE) hit = hit_sphere(ob->f, ray, t_min, closest, tmp);
    else if (type == OBJ_QUAD) hit = hit_quad(ob->f, ray, t_min, closest, tmp);
    else if (type == OBJ_CUBE) hit = hit_cube(ob->f, ray, t_min, closest, tmp);
    else if (type == OBJ_MESH) hit = hit_mesh<COUNT>(ob->f, sc.meshes[ob->mesh], ray, t_min, closest, tmp, ctr);
    else hit = hit_plane(ob->f, ray, t_min, closest, tmp);
    if (hit) {
      any = true;
      closest = tmp.t;
      best = tmp;
      best.object = i;
      best.material = ob->material;
    }
  }
  return any;
}

}  // namespace pt
