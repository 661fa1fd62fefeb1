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


# ----------------------------------------------------------------------------- seeded synthetic scenes


def random_scene(seed, n_tris=40, n_spheres=4, n_lights=2, max_depth=3, width=96, height=64, flat_fraction=0.3,
                 camera_inside_sphere=False, zero_specular=False):
    """A seeded triangle/sphere soup with the hard cases mixed in: axis-aligned (zero-thickness) triangles, triangles
    sharing edges and vertices (exact-t ties), a degenerate triangle, mirrors, lights inside geometry."""
    rng = np.random.default_rng(seed)
    verts, tris = [], []

    def add_vertex(p):
        verts.append([float(np.float32(x)) for x in p])
        return len(verts)

    n_mat = 5
    for i in range(n_tris):
        c = rng.uniform(-4, 4, 3) + np.array([0, 0, -10.0])
        a, b, d = c + rng.uniform(-1.5, 1.5, 3), c + rng.uniform(-1.5, 1.5, 3), c + rng.uniform(-1.5, 1.5, 3)
        if rng.random() < flat_fraction:  # axis-aligned: flat bounding box
            ax = int(rng.integers(0, 3))
            v = np.round(c[ax] * 4) / 4
            a[ax] = b[ax] = d[ax] = v
        ia, ib, ic = add_vertex(a), add_vertex(b), add_vertex(d)
        tris.append([ia, ib, ic, int(rng.integers(1, n_mat + 1))])
        if rng.random() < 0.4:  # a neighbour sharing the edge (b, d): rays through the edge tie exactly
            e = c + rng.uniform(-1.5, 1.5, 3)
            tris.append([ib, add_vertex(e), ic, int(rng.integers(1, n_mat + 1))])
    # a big floor quad (two coplanar triangles) and a degenerate triangle
    f = [add_vertex(p) for p in ([-8, -4, -2], [8, -4, -2], [8, -4, -18], [-8, -4, -18])]
    tris += [[f[0], f[1], f[2], 1], [f[2], f[3], f[0], 1], [f[0], f[0], f[1], 2]]
    sph_ids, sph_r = [], []
    for i in range(n_spheres):
        c = rng.uniform(-3, 3, 3) + np.array([0, 0, -9.0])
        sph_ids.append([int(rng.integers(1, n_mat + 1)), add_vertex(c)])
        sph_r.append(float(rng.uniform(0.4, 1.6)))
    if camera_inside_sphere:
        sph_ids.append([3, add_vertex([0.1, 0.0, 0.3])])
        sph_r.append(2.5)
    m13 = np.zeros((n_mat, 13), np.float32)
    m13[:, 0:3] = rng.uniform(0, 1, (n_mat, 3))
    m13[:, 3:6] = rng.uniform(0, 1, (n_mat, 3))
    m13[:, 6:9] = rng.uniform(0, 1, (n_mat, 3))
    m13[:, 9:12] = rng.uniform(0, 0.9, (n_mat, 3))
    m13[:, 12] = rng.choice([1, 2, 3, 10, 50, 100, 2.5], n_mat)
    mirror = (rng.random(n_mat) < 0.5).astype(np.int32)
    if zero_specular:  # materials whose specular term is exactly zero (the kernel skips computing it), one written as -0
        m13[0:3, 6:9] = 0.0
        m13[1, 6] = -0.0
        m13[2, 12] = 0.0  # pow(x, 0) == 1: still a zero term
        m13[3, 6:9] = [0.0, 0.5, 0.0]  # not ALL zero: the full path
    lights = np.zeros((n_lights, 6), np.float32)
    lights[:, 0:3] = rng.uniform(-6, 6, (n_lights, 3)) + np.array([0, 4, -6.0])
    lights[:, 3:6] = rng.uniform(200, 2000, (n_lights, 3))
    cam = RtCamera(RtVec3(0.1, 0.2, 0.5), RtVec3(0.02, -0.05, -1.0), RtVec3(0.0, 1.0, 0.05), -1.0, 1.0, -0.66, 0.66, 1.0, width, height)
    return Scene(np.array(verts, np.float32), np.array(tris, np.int32), np.array(sph_ids, np.int32).reshape(-1, 2), np.array(sph_r, np.float32),
                 m13, mirror, lights, [20.0, 25.0, 30.0], float(np.float32(rng.choice([1e-3, 1e-4, 1e-2]))), [10, 20, 30], max_depth,
                 [(cam, f"random_{seed}.ppm")])


def tessellated_scene(name, levels):
    """A shipped scene with every triangle cut into 4^levels coplanar sub-triangles (unshared vertices): the same
    picture from ~10^6 primitives, full of exact-t ties on the new shared edges — the large-scene stress for the GPU
    builders (multi-CTA PLOC, traversal stack depth) and for the tie ranks."""
    base = golden_scene(name)
    n = 1 << levels
    v = base.vertices.astype(np.float64)
    tri = base.triangles
    a, b, c = v[tri[:, 0] - 1], v[tri[:, 1] - 1], v[tri[:, 2] - 1]
    # barycentric lattice: point (i, j) = a + (b - a) i / n + (c - a) j / n, i + j <= n
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    keep = (ii + jj) <= n
    li, lj = ii[keep], jj[keep]
    index = -np.ones((n + 1, n + 1), np.int64)
    index[li, lj] = np.arange(len(li))
    pts = (a[:, None, :] + (b - a)[:, None, :] * (li / n)[None, :, None] + (c - a)[:, None, :] * (lj / n)[None, :, None]).astype(np.float32)
    per = pts.shape[1]
    small = []
    for i in range(n):
        for j in range(n - i):
            small.append([index[i, j], index[i + 1, j], index[i, j + 1]])
            if i + j + 1 < n:
                small.append([index[i + 1, j], index[i + 1, j + 1], index[i, j + 1]])
    small = np.array(small, np.int64)  # [4^levels, 3] lattice indices
    nt = len(tri)
    ids = (np.arange(nt)[:, None, None] * per + small[None, :, :] + 1).reshape(-1, 3)
    mats = np.repeat(tri[:, 3], len(small))
    new_tris = np.concatenate([ids, mats[:, None]], axis=1).astype(np.int32)
    # spheres keep their centres: append the original vertices behind the lattice points
    verts = np.concatenate([pts.reshape(-1, 3), base.vertices], axis=0)
    off = nt * per
    sph_ids = np.stack([base.spheres["material_id"], base.spheres["center_vertex_id"] + off], axis=1) if len(base.spheres) else np.zeros((0, 2), np.int32)
    d = base.desc
    return Scene(verts, new_tris, sph_ids, base.spheres["radius"] if len(base.spheres) else np.zeros(0, np.float32), base.materials["f"],
                 base.materials["is_mirror"], base.lights, (d.ambient_light.x, d.ambient_light.y, d.ambient_light.z), d.shadow_ray_epsilon,
                 list(d.background), d.max_recursion_depth, base.cameras)


def scene_to_xml(sc, path):
    """Writes a Scene in the reference's XML grammar (floats with 9 significant digits round-trip through >>)."""
    d = sc.desc

    def v(a):
        return " ".join(repr(float(np.float32(x))) if not float(x).is_integer() else str(int(x)) for x in a)

    def f32(x):
        return np.format_float_scientific(np.float32(x), unique=True)
    out = ["<Scene>", f"<BackgroundColor>{d.background[0]} {d.background[1]} {d.background[2]}</BackgroundColor>",
           f"<ShadowRayEpsilon>{f32(d.shadow_ray_epsilon)}</ShadowRayEpsilon>", f"<MaxRecursionDepth>{d.max_recursion_depth}</MaxRecursionDepth>", "<Cameras>"]
    for cam, name in sc.cameras:
        out += ["<Camera>", f"<Position>{f32(cam.position.x)} {f32(cam.position.y)} {f32(cam.position.z)}</Position>",
                f"<Gaze>{f32(cam.gaze.x)} {f32(cam.gaze.y)} {f32(cam.gaze.z)}</Gaze>", f"<Up>{f32(cam.up.x)} {f32(cam.up.y)} {f32(cam.up.z)}</Up>",
                f"<NearPlane>{f32(cam.l)} {f32(cam.r)} {f32(cam.b)} {f32(cam.t)}</NearPlane>", f"<NearDistance>{f32(cam.near_distance)}</NearDistance>",
                f"<ImageResolution>{cam.image_width} {cam.image_height}</ImageResolution>", f"<ImageName>{name}</ImageName>", "</Camera>"]
    out += ["</Cameras>", "<Lights>", f"<AmbientLight>{f32(d.ambient_light.x)} {f32(d.ambient_light.y)} {f32(d.ambient_light.z)}</AmbientLight>"]
    for l in sc.lights:
        out += ["<PointLight>", "<Position>" + " ".join(f32(x) for x in l[0:3]) + "</Position>",
                "<Intensity>" + " ".join(f32(x) for x in l[3:6]) + "</Intensity>", "</PointLight>"]
    out += ["</Lights>", "<Materials>"]
    for m in sc.materials:
        fm = m["f"]
        out += ['<Material type="mirror">' if m["is_mirror"] else "<Material>",
                "<AmbientReflectance>" + " ".join(f32(x) for x in fm[0:3]) + "</AmbientReflectance>",
                "<DiffuseReflectance>" + " ".join(f32(x) for x in fm[3:6]) + "</DiffuseReflectance>",
                "<SpecularReflectance>" + " ".join(f32(x) for x in fm[6:9]) + "</SpecularReflectance>",
                "<MirrorReflectance>" + " ".join(f32(x) for x in fm[9:12]) + "</MirrorReflectance>",
                f"<PhongExponent>{f32(fm[12])}</PhongExponent>", "</Material>"]
    out += ["</Materials>", "<VertexData>"] + [" ".join(f32(x) for x in p) for p in sc.vertices] + ["</VertexData>", "<Objects>"]
    for t in sc.triangles:  # every triangle as a <Triangle>: the flat list order is then the file order
        out += ["<Triangle>", f"<Material>{t[3]}</Material>", f"<Indices>{t[0]} {t[1]} {t[2]}</Indices>", "</Triangle>"]
    for s in sc.spheres:
        out += ["<Sphere>", f"<Material>{s['material_id']}</Material>", f"<Center>{s['center_vertex_id']}</Center>",
                f"<Radius>{f32(s['radius'])}</Radius>", "</Sphere>"]
    out += ["</Objects>", "</Scene>"]
    with open(path, "w") as fo:
        fo.write("\n".join(out))
