"""ctypes bindings of the two TEST-ONLY libraries:

  oracle/liboracle.so          CPU restatement of the reference (the checker)
  tests/hostsim/libhostsim.so  the product's device headers compiled by g++ (logic check without a GPU)

Both take the same scene description as the product (a `Scene` of the package), so one description feeds the
checker and the thing being checked.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_F = C.POINTER(C.c_float)
_vp = C.c_void_p


class OrcSettings(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("seed", C.c_uint64), ("rng_mode", C.c_int32), ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
                ("threads", C.c_int32)]


class OrcStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("seconds", C.c_double)]


RNG_CHACHA, RNG_PHILOX = 0, 1


def _ensure(path, target):
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", ROOT, target])
    return path


_oracle = None
_sim = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        L = C.CDLL(_ensure(os.path.join(ROOT, "oracle", "liboracle.so"), "oracle/liboracle.so"))
        L.orc_scene_create.restype = _vp
        L.orc_scene_destroy.argtypes = [_vp]
        L.orc_scene_destroy.restype = None
        L.orc_scene_add_material.argtypes = [_vp, _vp]
        L.orc_scene_add_sphere.argtypes = [_vp, _F, C.c_float, C.c_int]
        L.orc_scene_add_plane.argtypes = [_vp, _F, _F, C.c_int]
        L.orc_scene_add_quad.argtypes = [_vp, _F, _F, _F, _F, C.c_float, C.c_float, C.c_float, C.c_int]
        L.orc_scene_add_cube.argtypes = [_vp, _F, _F, C.c_int]
        L.orc_scene_add_mesh.argtypes = [_vp, _F, C.c_int64, _F, _F, C.c_int]
        L.orc_scene_set_sky_hdr.argtypes = [_vp, _F, C.c_int, C.c_int]
        L.orc_intersect.argtypes = [_vp, _F, _F, C.c_int64, C.c_float, C.c_float, _vp]
        L.orc_render.argtypes = [_vp, _vp, C.POINTER(OrcSettings), _F, C.POINTER(OrcStats)]
        L.orc_resolve_u32.argtypes = [_F, C.c_int64, _vp]
        L.orc_resolve_u32.restype = None
        L.orc_camera_get_ray.argtypes = [_vp, C.c_float, C.c_float, _F, _F]
        L.orc_camera_get_ray.restype = None
        L.orc_scatter.argtypes = [_vp, _F, _F, _F, C.c_int, _F, _F, _F, _F, _F]
        L.orc_mesh_bvh_info.argtypes = [_vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), _vp, _vp]
        L.orc_chacha_stream.argtypes = [C.c_uint64, _vp, C.c_int]
        L.orc_chacha_stream.restype = None
        L.orc_philox4x32_10.argtypes = [_vp, _vp, _vp]
        L.orc_philox4x32_10.restype = None
        L.orc_chacha_block.argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_int, _vp]
        L.orc_chacha_block.restype = None
        _oracle = L
    return _oracle


def sim_lib():
    global _sim
    if _sim is None:
        L = C.CDLL(_ensure(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"), "tests/hostsim/libhostsim.so"))
        L.sim_last_error.restype = C.c_char_p
        L.sim_scene_create.restype = _vp
        L.sim_scene_destroy.argtypes = [_vp]
        L.sim_scene_destroy.restype = None
        L.sim_scene_add_material.argtypes = [_vp, _vp]
        L.sim_scene_add_sphere.argtypes = [_vp, _F, C.c_float, C.c_int]
        L.sim_scene_add_plane.argtypes = [_vp, _F, _F, C.c_int]
        L.sim_scene_add_quad.argtypes = [_vp, _F, _F, _F, _F, C.c_float, C.c_float, C.c_float, C.c_int]
        L.sim_scene_add_cube.argtypes = [_vp, _F, _F, C.c_int]
        L.sim_scene_add_mesh.argtypes = [_vp, _F, C.c_int64, _F, _F, C.c_int]
        L.sim_scene_set_sky_hdr.argtypes = [_vp, _F, C.c_int32, C.c_int32]
        L.sim_scene_commit.argtypes = [_vp, C.c_int]
        L.sim_scene_mesh_info.argtypes = [_vp, C.c_int, _vp, _vp, _vp]
        L.sim_intersect.argtypes = [_vp, _F, _F, C.c_int64, C.c_float, C.c_float, _vp, _vp]
        L.sim_primary_rays.argtypes = [_vp, _vp, _vp, C.c_int32, _F, _F]
        L.sim_scatter.argtypes = [_vp, C.c_int, _F, _F, _F, _vp, _F, C.c_int64, _vp, _F, _F, _F, _F]
        L.sim_philox.argtypes = [_vp, _vp, _vp, _vp]
        L.sim_resolve_u32.argtypes = [_vp, _F, C.c_int64, C.c_float, _vp]
        _sim = L
    return _sim


def _fp(a):
    return a.ctypes.data_as(_F)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


def _feed(L, prefix, handle, scene):
    """Replay a package `Scene` into a library with the add_* vocabulary (object_list order preserved)."""
    g = lambda n: getattr(L, prefix + n)  # noqa: E731
    for m in scene.materials:
        assert g("scene_add_material")(handle, C.addressof(m)) >= 0
    for o in scene.objects:
        if o.type == 0:
            r = g("scene_add_sphere")(handle, o.center, o.radius, o.material)
        elif o.type == 1:
            r = g("scene_add_plane")(handle, o.p1, o.normal, o.material)
        elif o.type == 2:
            r = g("scene_add_quad")(handle, o.base, o.edge0, o.edge1, o.normal, o.d, o.inv_edge0_len_sq, o.inv_edge1_len_sq, o.material)
        elif o.type == 3:
            r = g("scene_add_cube")(handle, o.o2w, o.w2o, o.material)
        else:
            t = scene.mesh(o.mesh)
            r = g("scene_add_mesh")(handle, _fp(t), len(t), o.o2w, o.w2o, o.material)
        assert r >= 0
    sky = scene.sky
    if sky is not None:
        assert g("scene_set_sky_hdr")(handle, _fp(sky), sky.shape[1], sky.shape[0]) == 0


class OracleScene:
    def __init__(self, scene):
        self.L = oracle_lib()
        self.h = self.L.orc_scene_create()
        self.scene = scene
        _feed(self.L, "orc_", self.h, scene)

    def __del__(self):
        try:
            self.L.orc_scene_destroy(self.h)
        except Exception:
            pass

    def intersect(self, pt, origins, dirs, t_min=1e-4, t_max=float("inf")):
        o, d = _f32(origins, (-1, 3)), _f32(dirs, (-1, 3))
        out = np.zeros(len(o), pt.HIT_DTYPE)
        assert self.L.orc_intersect(self.h, _fp(o), _fp(d), len(o), t_min, t_max, out.ctypes.data) == 0
        return out

    def render(self, camera, width, height, spp, max_depth, rng_mode=RNG_PHILOX, seed=0, sample_begin=0, sample_end=0,
               threads=0):
        st = OrcSettings(width=width, height=height, spp=spp, max_depth=max_depth, seed=seed, rng_mode=rng_mode,
                         sample_begin=sample_begin, sample_end=sample_end, threads=threads)
        img = np.zeros((height, width, 3), np.float32)
        stats = OrcStats()
        assert self.L.orc_render(self.h, C.addressof(camera), C.byref(st), _fp(img), C.byref(stats)) == 0
        return img, stats

    def mesh_bvh_info(self, obj, n_tris):
        nodes, leaves, depth = C.c_int64(), C.c_int64(), C.c_int32()
        dead = np.zeros(n_tris, np.uint8)
        order = np.zeros(n_tris, np.int32)
        assert self.L.orc_mesh_bvh_info(self.h, obj, C.byref(nodes), C.byref(leaves), C.byref(depth), dead.ctypes.data,
                                        order.ctypes.data) == 0
        return nodes.value, leaves.value, depth.value, dead, order


def oracle_scatter(material, ray_dirs, positions, normals, front_face, u4):
    L = oracle_lib()
    d, p, nn, u = _f32(ray_dirs, (-1, 3)), _f32(positions, (-1, 3)), _f32(normals, (-1, 3)), _f32(u4, (-1, 4))
    n = len(d)
    sc = np.zeros(n, np.int32)
    oo, od, att, em = (np.zeros((n, 3), np.float32) for _ in range(4))
    for i in range(n):
        sc[i] = L.orc_scatter(C.addressof(material), _fp(d[i]), _fp(p[i]), _fp(nn[i]), int(front_face[i]), _fp(u[i]),
                              _fp(oo[i]), _fp(od[i]), _fp(att[i]), _fp(em[i]))
    return sc, oo, od, att, em


def oracle_resolve(img):
    a = _f32(img).reshape(-1, 3)
    out = np.zeros(len(a), np.uint32)
    oracle_lib().orc_resolve_u32(_fp(a), len(a), out.ctypes.data)
    return out


def oracle_get_ray(camera, u, v):
    o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    oracle_lib().orc_camera_get_ray(C.addressof(camera), u, v, _fp(o), _fp(d))
    return o, d


def oracle_philox(ctr, key):
    c, k, out = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    oracle_lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


class SimScene:
    """Same calls as the package's CoreScene, served by the g++ build of the device headers."""

    def __init__(self, scene):
        self.L = sim_lib()
        self.h = self.L.sim_scene_create()
        self.scene = scene
        _feed(self.L, "sim_", self.h, scene)
        assert self.L.sim_scene_commit(self.h, 0) == 0, self.L.sim_last_error()

    def __del__(self):
        try:
            self.L.sim_scene_destroy(self.h)
        except Exception:
            pass

    def mesh_info(self, pt, obj):
        info = pt.MeshInfo()
        assert self.L.sim_scene_mesh_info(self.h, obj, C.addressof(info), None, None) == 0
        dead, order = np.zeros(info.triangles, np.uint8), np.zeros(info.triangles, np.int32)
        assert self.L.sim_scene_mesh_info(self.h, obj, C.addressof(info), dead.ctypes.data, order.ctypes.data) == 0
        return info, dead, order

    def intersect(self, pt, origins, dirs, t_min=1e-4, t_max=float("inf")):
        o, d = _f32(origins, (-1, 3)), _f32(dirs, (-1, 3))
        out = np.zeros(len(o), pt.HIT_DTYPE)
        st = pt.Stats()
        assert self.L.sim_intersect(self.h, _fp(o), _fp(d), len(o), t_min, t_max, out.ctypes.data, C.addressof(st)) == 0
        return out, st

    def primary_rays(self, camera, settings, sample):
        n = settings.width * settings.height
        o, d = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        assert self.L.sim_primary_rays(self.h, C.addressof(camera), C.addressof(settings), sample, _fp(o), _fp(d)) == 0
        return o, d

    def scatter(self, material, ray_dirs, positions, normals, front_face, u4):
        d, p, nn, u = _f32(ray_dirs, (-1, 3)), _f32(positions, (-1, 3)), _f32(normals, (-1, 3)), _f32(u4, (-1, 4))
        ff = np.ascontiguousarray(front_face, dtype=np.int32)
        n = len(d)
        sc = np.zeros(n, np.int32)
        oo, od, att, em = (np.zeros((n, 3), np.float32) for _ in range(4))
        assert self.L.sim_scatter(self.h, material, _fp(d), _fp(p), _fp(nn), ff.ctypes.data, _fp(u), n, sc.ctypes.data, _fp(oo),
                                  _fp(od), _fp(att), _fp(em)) == 0
        return sc, oo, od, att, em

    def philox(self, ctr, key):
        c, k, out = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
        self.L.sim_philox(self.h, c.ctypes.data, k.ctypes.data, out.ctypes.data)
        return out

    def resolve_u32(self, rgb, scale=1.0):
        a = _f32(rgb).reshape(-1, 3)
        out = np.zeros(len(a), np.uint32)
        self.L.sim_resolve_u32(self.h, _fp(a), len(a), scale, out.ctypes.data)
        return out


def compare_hits(got, want, rel_t=1e-5):
    """The parity bar of the north star: ids bit-exact, t within 1e-5 relative.  Returns a dict of mismatch counts."""
    id_bad = (got["object"] != want["object"]) | (got["triangle"] != want["triangle"])
    both = (got["object"] >= 0) & (want["object"] >= 0) & ~id_bad
    t_bad = np.zeros(len(got), bool)
    t_bad[both] = np.abs(got["t"][both] - want["t"][both]) > rel_t * np.abs(want["t"][both])
    exact = both & (got["t"] == want["t"]) & (got["position"] == want["position"]).all(1) & \
        (got["normal"] == want["normal"]).all(1) & (got["front_face"] == want["front_face"])
    return {"n": len(got), "id_mismatch": int(id_bad.sum()), "t_mismatch": int(t_bad.sum()),
            "hits": int((want["object"] >= 0).sum()), "bit_exact_records": int(exact.sum()), "both_hit": int(both.sum()),
            "id_bad_idx": np.nonzero(id_bad)[0]}
