"""Test/bench plumbing: binds the CHECKERS (oracle, reference) and the golden fixtures.

  * the PRODUCT lives in raytracer-ceng477-graphics-hw-1_b200/rt_b200.py (re-exported here)
  * the ORACLE   : oracle/liboracle.so (plain-C restatement; checker only)
  * the REFERENCE: oracle/_ref/libref.so (unmodified reference behind oracle/ref_shim.cpp; checker /
                   CPU baseline only; built in the authoring container, shipped to the GPU box)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raytracer-ceng477-graphics-hw-1_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

import rt_b200  # noqa: E402
from rt_b200 import (RayTracer, RtBuildOptions, RtCamera, RtMaterial, RtPointLight, RtSceneDesc, RtSceneInfo,  # noqa: E402,F401
                     RtSphere, RtStats, RtTriangle, RtVec3, Scene, host_lib, load_scene_xml, write_ppm)


class OrStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("reflection_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shadow_occluded", C.c_uint64), ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64)]

    @property
    def total_rays(self):
        return self.primary_rays + self.reflection_rays + self.shadow_rays


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


# ----------------------------------------------------------------------------- golden fixtures

_manifest = None
_scene_dir = None


def manifest():
    global _manifest
    if _manifest is None:
        import json
        with open(os.path.join(GOLDEN, "manifest.json")) as f:
            _manifest = json.load(f)
    return _manifest


def golden_scene_path(name):
    """Decompresses tests/golden/scenes/<name>.xml.xz into a per-process temp dir, returns the .xml path."""
    global _scene_dir
    import lzma
    import tempfile
    if _scene_dir is None:
        _scene_dir = tempfile.mkdtemp(prefix="rtb200_scenes_")
    out = os.path.join(_scene_dir, name + ".xml")
    if not os.path.exists(out):
        with lzma.open(os.path.join(GOLDEN, "scenes", name + ".xml.xz")) as f, open(out + ".tmp", "wb") as g:
            g.write(f.read())
        os.replace(out + ".tmp", out)
    return out


_scene_cache = {}


def golden_scene(name):
    if name not in _scene_cache:
        _scene_cache[name] = load_scene_xml(golden_scene_path(name))
    return _scene_cache[name]


def golden_image(key):
    import lzma
    m = manifest()["images"][key]
    with lzma.open(os.path.join(GOLDEN, "images", key + ".rgb.xz")) as f:
        raw = f.read()
    rows = len(m["rows"]) if "rows" in m else m["height"]
    return np.frombuffer(raw, np.uint8).reshape(rows, m["width"], 3), m
