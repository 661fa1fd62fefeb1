"""ctypes plumbing shared by the tests, bench.py and __graft_entry__.py.

Three libraries are bound here:
  * the PRODUCT  : raytracer-ceng477-graphics-hw-1_b200/libwhitted_b200.so (CUDA, C-ABI of include/rt_b200.h)
                   and .../libwhitted_host.so (XML scene reader + PPM writer, host only)
  * the ORACLE   : oracle/liboracle.so (plain-C restatement; checker only)
  * the REFERENCE: oracle/_ref/libref.so (unmodified reference behind oracle/ref_shim.cpp; checker /
                   CPU baseline only; built in the authoring container, shipped to the GPU box)
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raytracer-ceng477-graphics-hw-1_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class RtVec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class RtMaterial(C.Structure):
    _fields_ = [("ambient", RtVec3), ("diffuse", RtVec3), ("specular", RtVec3), ("mirror", RtVec3),
                ("phong_exponent", C.c_float), ("is_mirror", C.c_int32)]


class RtPointLight(C.Structure):
    _fields_ = [("position", RtVec3), ("intensity", RtVec3)]


class RtTriangle(C.Structure):
    _fields_ = [("v0_id", C.c_int32), ("v1_id", C.c_int32), ("v2_id", C.c_int32), ("material_id", C.c_int32)]


class RtSphere(C.Structure):
    _fields_ = [("material_id", C.c_int32), ("center_vertex_id", C.c_int32), ("radius", C.c_float)]


class RtSceneDesc(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("n_vertices", C.c_int32),
                ("triangles", C.c_void_p), ("n_triangles", C.c_int32),
                ("spheres", C.c_void_p), ("n_spheres", C.c_int32),
                ("materials", C.c_void_p), ("n_materials", C.c_int32),
                ("lights", C.c_void_p), ("n_lights", C.c_int32),
                ("ambient_light", RtVec3), ("background", C.c_int32 * 3),
                ("shadow_ray_epsilon", C.c_float), ("max_recursion_depth", C.c_int32)]


class RtCamera(C.Structure):
    _fields_ = [("position", RtVec3), ("gaze", RtVec3), ("up", RtVec3),
                ("l", C.c_float), ("r", C.c_float), ("b", C.c_float), ("t", C.c_float),
                ("near_distance", C.c_float), ("image_width", C.c_int32), ("image_height", C.c_int32)]


class RtBuildOptions(C.Structure):
    _fields_ = [("builder", C.c_int32), ("brute_force", C.c_int32), ("reserved", C.c_int32 * 6)]


class RtStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("reflection_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shadow_occluded", C.c_uint64), ("ms_render", C.c_float), ("ms_d2h", C.c_float),
                ("ms_total", C.c_float), ("n_launches", C.c_int32), ("reserved", C.c_int32 * 3)]

    @property
    def total_rays(self):
        return self.primary_rays + self.reflection_rays + self.shadow_rays


class RtSceneInfo(C.Structure):
    _fields_ = [("n_triangles", C.c_int32), ("n_spheres", C.c_int32), ("bvh_nodes", C.c_int32),
                ("bvh_max_depth", C.c_int32), ("ref_tree_nodes", C.c_int32), ("ref_tree_leaves", C.c_int32),
                ("ref_tree_max_leaf", C.c_int32), ("ref_tree_max_depth", C.c_int32),
                ("ms_build_host", C.c_float), ("ms_build_device", C.c_float), ("bvh_sah_cost", C.c_float),
                ("builder", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32 * 3)]


class OrStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("reflection_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shadow_occluded", C.c_uint64), ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64)]

    @property
    def total_rays(self):
        return self.primary_rays + self.reflection_rays + self.shadow_rays


class Scene:
    """Flat scene arrays (numpy) + the RtSceneDesc pointing at them + cameras."""

    def __init__(self, vertices, triangles, sphere_ids, sphere_radius, materials13, is_mirror, lights6,
                 ambient, eps, background, max_depth, cameras):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 4)
        ns = len(sphere_radius)
        self.spheres = np.zeros(ns, dtype=np.dtype([("material_id", "<i4"), ("center_vertex_id", "<i4"), ("radius", "<f4")]))
        if ns:
            ids = np.asarray(sphere_ids, dtype=np.int32).reshape(-1, 2)
            self.spheres["material_id"] = ids[:, 0]
            self.spheres["center_vertex_id"] = ids[:, 1]
            self.spheres["radius"] = np.asarray(sphere_radius, dtype=np.float32)
        m13 = np.asarray(materials13, dtype=np.float32).reshape(-1, 13)
        self.materials = np.zeros(len(m13), dtype=np.dtype([("f", "<f4", 13), ("is_mirror", "<i4")]))
        self.materials["f"] = m13
        self.materials["is_mirror"] = np.asarray(is_mirror, dtype=np.int32)
        self.lights = np.ascontiguousarray(lights6, dtype=np.float32).reshape(-1, 6)
        self.cameras = cameras  # list of (RtCamera, name)
        d = RtSceneDesc()
        d.vertices = self.vertices.ctypes.data
        d.n_vertices = len(self.vertices)
        d.triangles = self.triangles.ctypes.data
        d.n_triangles = len(self.triangles)
        d.spheres = self.spheres.ctypes.data
        d.n_spheres = ns
        d.materials = self.materials.ctypes.data
        d.n_materials = len(self.materials)
        d.lights = self.lights.ctypes.data
        d.n_lights = len(self.lights)
        d.ambient_light = RtVec3(*[float(a) for a in ambient])
        d.background = (C.c_int32 * 3)(*[int(b) for b in background])
        d.shadow_ray_epsilon = float(eps)
        d.max_recursion_depth = int(max_depth)
        self.desc = d

    def camera(self, name_or_index=0, width=None, height=None):
        if isinstance(name_or_index, int):
            cam, name = self.cameras[name_or_index]
        else:
            cam, name = next((c, n) for c, n in self.cameras if n == name_or_index or n == name_or_index + ".ppm")
        out = RtCamera.from_buffer_copy(cam)
        if width:
            out.image_width = width
        if height:
            out.image_height = height
        return out

    def digest(self):
        """sha256 over every parsed value (loader parity)."""
        import hashlib
        h = hashlib.sha256()
        for a in (self.vertices, self.triangles, self.spheres, self.materials, self.lights):
            h.update(a.tobytes())
        d = self.desc
        h.update(np.array([d.ambient_light.x, d.ambient_light.y, d.ambient_light.z, d.shadow_ray_epsilon], dtype=np.float32).tobytes())
        h.update(np.array(list(d.background) + [d.max_recursion_depth], dtype=np.int32).tobytes())
        for cam, name in self.cameras:
            h.update(bytes(cam))
            h.update(name.encode())
        return h.hexdigest()


# ----------------------------------------------------------------------------- reference (libref.so)

_ref = None


def ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref.so"))


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_last_error.restype = C.c_char_p
        L.ref_build_seconds.restype = C.c_double
        L.ref_build_seconds.argtypes = [C.c_void_p]
        for fn in ("ref_counts", "ref_copy_vertices", "ref_copy_triangles", "ref_bvh_stats"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p]
        for fn in ("ref_copy_spheres", "ref_copy_materials", "ref_copy_globals"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_copy_lights.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_copy_camera.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ref_render.restype = C.c_double
        L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_time_rows.restype = C.c_double
        L.ref_time_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                    C.c_int, C.c_int, C.c_void_p]
        L.ref_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_hardware_threads.restype = C.c_int
        _ref = L
    return _ref


class RefScene:
    """The unmodified reference: its loader, its BVH, its renderer."""

    def __init__(self, xml_path):
        self.L = ref_lib()
        self.h = self.L.ref_open(xml_path.encode())
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        cnt = (C.c_int * 8)()
        self.L.ref_counts(self.h, cnt)
        self.counts = list(cnt)

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None

    def to_scene(self):
        nv, ntri, nmesh, nfaces, ns, nm, nl, nc = self.counts
        v = np.zeros((nv, 3), np.float32)
        self.L.ref_copy_vertices(self.h, v.ctypes.data)
        t = np.zeros((ntri + nfaces, 4), np.int32)
        self.L.ref_copy_triangles(self.h, t.ctypes.data)
        sid = np.zeros((ns, 2), np.int32)
        srad = np.zeros(ns, np.float32)
        self.L.ref_copy_spheres(self.h, sid.ctypes.data, srad.ctypes.data)
        m13 = np.zeros((nm, 13), np.float32)
        mir = np.zeros(nm, np.int32)
        self.L.ref_copy_materials(self.h, m13.ctypes.data, mir.ctypes.data)
        l6 = np.zeros((nl, 6), np.float32)
        self.L.ref_copy_lights(self.h, l6.ctypes.data)
        f4 = np.zeros(4, np.float32)
        i4 = np.zeros(4, np.int32)
        self.L.ref_copy_globals(self.h, f4.ctypes.data, i4.ctypes.data)
        cams = []
        for i in range(nc):
            f14 = np.zeros(14, np.float32)
            wh = np.zeros(2, np.int32)
            name = C.create_string_buffer(512)
            self.L.ref_copy_camera(self.h, i, f14.ctypes.data, wh.ctypes.data, name, 512)
            cam = RtCamera(RtVec3(*f14[0:3]), RtVec3(*f14[3:6]), RtVec3(*f14[6:9]), f14[9], f14[10], f14[11], f14[12],
                           f14[13], int(wh[0]), int(wh[1]))
            cams.append((cam, name.value.decode()))
        return Scene(v, t, sid, srad, m13, mir, l6, f4[:3], f4[3], i4[:3], i4[3], cams)

    def bvh_stats(self):
        o = (C.c_int * 4)()
        self.L.ref_bvh_stats(self.h, o)
        return list(o)

    def render(self, cam_idx=0, aa=1, width=0, height=0):
        f14 = np.zeros(14, np.float32)
        wh = np.zeros(2, np.int32)
        name = C.create_string_buffer(512)
        self.L.ref_copy_camera(self.h, cam_idx, f14.ctypes.data, wh.ctypes.data, name, 512)
        w = width or int(wh[0])
        h = height or int(wh[1])
        out = np.zeros((h, w, 3), np.uint8)
        secs = self.L.ref_render(self.h, cam_idx, aa, width, height, out.ctypes.data)
        if secs < 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        return out, secs

    def time_rows(self, cam_idx, aa, width, height, row0, row_stride, n_rows, threads=0, keep=False):
        rows = np.zeros((n_rows, (width) * aa, 3), np.uint8) if keep else None
        secs = self.L.ref_time_rows(self.h, cam_idx, aa, width, height, row0, row_stride, n_rows, threads,
                                    rows.ctypes.data if keep else None)
        return secs, rows


# ----------------------------------------------------------------------------- oracle (liboracle.so)

_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.or_scene_create.restype = C.c_void_p
        L.or_scene_create.argtypes = [C.POINTER(RtSceneDesc)]
        L.or_scene_destroy.argtypes = [C.c_void_p]
        L.or_bvh_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.or_visit_ranks.argtypes = [C.c_void_p, C.c_void_p]
        L.or_render.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_int, C.c_void_p, C.POINTER(OrStats), C.c_int]
        L.or_render_rows.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_int, C.c_longlong, C.c_longlong, C.c_int,
                                     C.c_void_p, C.POINTER(OrStats), C.c_int]
        L.or_specular_gate.argtypes = [C.c_float]
        _oracle = L
    return _oracle


class OracleScene:
    def __init__(self, scene):
        self.L = oracle_lib()
        self.scene = scene
        self.h = self.L.or_scene_create(C.byref(scene.desc))

    def close(self):
        if self.h:
            self.L.or_scene_destroy(self.h)
            self.h = None

    def bvh_stats(self):
        o = (C.c_int * 4)()
        self.L.or_bvh_stats(self.h, o)
        return list(o)

    def visit_ranks(self):
        n = self.scene.desc.n_triangles + self.scene.desc.n_spheres
        out = np.zeros((8, n), np.uint32)
        self.L.or_visit_ranks(self.h, out.ctypes.data)
        return out

    def render(self, cam, aa=1, threads=None):
        out = np.zeros((cam.image_height, cam.image_width, 3), np.uint8)
        st = OrStats()
        self.L.or_render(self.h, C.byref(cam), aa, out.ctypes.data, C.byref(st), threads or os.cpu_count() or 8)
        return out, st

    def render_rows(self, cam, aa, row0, row_stride, n_rows, threads=None, keep=True):
        rows = np.zeros((n_rows, cam.image_width * aa, 3), np.uint8) if keep else None
        st = OrStats()
        self.L.or_render_rows(self.h, C.byref(cam), aa, row0, row_stride, n_rows, rows.ctypes.data if keep else None,
                              C.byref(st), threads or os.cpu_count() or 8)
        return rows, st


# ----------------------------------------------------------------------------- image comparison


def diff_report(a, b):
    """The tolerance bookkeeping of SURVEY.md appendix C.4 on two HxWx3 uint8 images."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16)).max(axis=2)
    n = d.size
    return {"pixels": int(n), "equal": int((d == 0).sum()), "le1": int((d <= 1).sum()), "gt1": int((d > 1).sum()),
            "gt8": int((d > 8).sum()), "max": int(d.max()) if n else 0}


def within_tolerance(rep):
    """north_star: |delta| <= 1 per channel on >= 99.9 % of pixels, zero pixels off by more than 8."""
    return rep["gt8"] == 0 and rep["le1"] >= 0.999 * rep["pixels"]
