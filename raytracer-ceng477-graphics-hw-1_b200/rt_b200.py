"""rt_b200 — Python host-side binding of the B200-native Whitted hot path.

Mirrors the reference's host interface for the path (raytracer.cpp:327-383):

    reference (C++)                              here
    ------------------------------------------   ---------------------------------------------
    parser::Scene scene; scene.loadFromXml(p)    scene = load_scene_xml(p)
    RayTracer rayTracer(scene)                   tracer = RayTracer(scene)          # BVH build, upload
    Image img = rayTracer.render(camera)         img = tracer.render(camera, aa)    # HxWx3 uint8
    ImageProcessor::downSample(img, w, h, f)     (fused into render: aa=f)
    write_ppm(name, img, w, h)                   write_ppm(name, img)

Everything heavy happens behind the C-ABI of include/rt_b200.h in libwhitted_b200.so (CUDA,
sm_100a).  There is NO CPU fallback: if the library is missing or no GPU is usable, RayTracer raises.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)

RT_BUILD_DEFAULT, RT_BUILD_LBVH_GPU, RT_BUILD_SAH_HOST, RT_BUILD_PLOC_GPU, RT_BUILD_AUTO, RT_BUILD_SAH_GPU = 0, 1, 2, 3, 4, 5


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
    _fields_ = [("builder", C.c_int32), ("brute_force", C.c_int32), ("no_exact_culling", C.c_int32),
                ("refill_threshold", C.c_int32), ("ploc_radius", C.c_int32), ("ploc_leaf_cost", C.c_float),
                ("force_replay", C.c_int32), ("max_ctas_per_sm", C.c_int32),
                ("reinsert_rounds", C.c_int32), ("reinsert_accept", C.c_float)]


class RtStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("reflection_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shadow_occluded", C.c_uint64), ("replayed_closest", C.c_uint64), ("replayed_any", C.c_uint64),
                ("ms_render", C.c_float), ("ms_d2h", C.c_float),
                ("ms_total", C.c_float), ("n_launches", C.c_int32), ("reserved", C.c_int32 * 3)]

    @property
    def total_rays(self):
        return self.primary_rays + self.reflection_rays + self.shadow_rays


class RtSceneInfo(C.Structure):
    _fields_ = [("n_triangles", C.c_int32), ("n_spheres", C.c_int32), ("bvh_nodes", C.c_int32),
                ("bvh_max_depth", C.c_int32), ("ref_tree_nodes", C.c_int32), ("ref_tree_leaves", C.c_int32),
                ("ref_tree_max_leaf", C.c_int32), ("ref_tree_max_depth", C.c_int32),
                ("ms_build_host", C.c_float), ("ms_build_device", C.c_float), ("bvh_sah_cost", C.c_float),
                ("builder", C.c_int32), ("device", C.c_int32), ("sah_cost_ploc", C.c_float), ("sah_cost_sah", C.c_float),
                ("ms_create_wall", C.c_float), ("reinsert_cost_before", C.c_float), ("reinsert_cost_after", C.c_float),
                ("reinsert_moves", C.c_int32), ("reinsert_rounds", C.c_int32), ("reinsert_accepted", C.c_int32)]


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


# ----------------------------------------------------------------------------- host library (XML, PPM)

_host = None


def host_lib():
    global _host
    if _host is None:
        path = os.path.join(PKG, "libwhitted_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.rth_last_error.restype = C.c_char_p
        L.rth_scene_load_xml.restype = C.c_void_p
        L.rth_scene_load_xml.argtypes = [C.c_char_p]
        L.rth_scene_free.argtypes = [C.c_void_p]
        L.rth_scene_desc.restype = C.POINTER(RtSceneDesc)
        L.rth_scene_desc.argtypes = [C.c_void_p]
        L.rth_scene_num_cameras.argtypes = [C.c_void_p]
        L.rth_scene_camera.argtypes = [C.c_void_p, C.c_int, C.POINTER(RtCamera), C.c_char_p, C.c_int]
        L.rth_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        _host = L
    return _host


def _np_from(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


def load_scene_xml(path):
    """Scene via the PRODUCT's loader (csrc/host/xml_scene.cpp)."""
    L = host_lib()
    h = L.rth_scene_load_xml(path.encode())
    if not h:
        raise RuntimeError(L.rth_last_error().decode())
    try:
        d = L.rth_scene_desc(h).contents
        v = _np_from(d.vertices, d.n_vertices * 3, np.float32).reshape(-1, 3)
        t = _np_from(d.triangles, d.n_triangles * 4, np.int32).reshape(-1, 4)
        sp = _np_from(d.spheres, d.n_spheres, np.dtype([("material_id", "<i4"), ("center_vertex_id", "<i4"), ("radius", "<f4")]))
        mt = _np_from(d.materials, d.n_materials, np.dtype([("f", "<f4", 13), ("is_mirror", "<i4")]))
        l6 = _np_from(d.lights, d.n_lights * 6, np.float32).reshape(-1, 6)
        cams = []
        for i in range(L.rth_scene_num_cameras(h)):
            cam = RtCamera()
            name = C.create_string_buffer(512)
            L.rth_scene_camera(h, i, C.byref(cam), name, 512)
            cams.append((cam, name.value.decode()))
        sid = np.stack([sp["material_id"], sp["center_vertex_id"]], axis=1) if len(sp) else np.zeros((0, 2), np.int32)
        return Scene(v, t, sid, sp["radius"] if len(sp) else np.zeros(0, np.float32), mt["f"] if len(mt) else np.zeros((0, 13)),
                     mt["is_mirror"] if len(mt) else np.zeros(0, np.int32), l6,
                     (d.ambient_light.x, d.ambient_light.y, d.ambient_light.z), d.shadow_ray_epsilon,
                     list(d.background), d.max_recursion_depth, cams)
    finally:
        L.rth_scene_free(h)


def write_ppm(path, img):
    L = host_lib()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if L.rth_write_ppm(path.encode(), img.ctypes.data, img.shape[1], img.shape[0]) != 0:
        raise RuntimeError(L.rth_last_error().decode())




# ----------------------------------------------------------------------------- CUDA library (C-ABI)

_cuda = None


class RtError(RuntimeError):
    pass


def cuda_lib():
    """libwhitted_b200.so; raises if it has not been built (no fallback)."""
    global _cuda
    if _cuda is None:
        path = os.environ.get("RT_B200_LIB") or os.path.join(PKG, "libwhitted_b200.so")  # env override: A/B kernel experiments
        if not os.path.exists(path):
            raise RtError(f"{path} is missing: build it with `make -C {PKG}` (or __graft_entry__.build())")
        L = C.CDLL(path)
        L.rt_last_error.restype = C.c_char_p
        L.rt_scene_create.argtypes = [C.POINTER(RtSceneDesc), C.POINTER(RtBuildOptions), C.POINTER(C.c_void_p)]
        L.rt_scene_destroy.argtypes = [C.c_void_p]
        L.rt_scene_destroy.restype = None
        L.rt_scene_info.argtypes = [C.c_void_p, C.POINTER(RtSceneInfo)]
        L.rt_render.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_int, C.c_void_p, C.POINTER(RtStats)]
        L.rt_render_async.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.rt_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(RtStats)]
        L.rt_host_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p)]
        L.rt_host_free.argtypes = [C.c_void_p]
        L.rt_band_height.argtypes = [C.POINTER(RtCamera), C.c_int, C.c_int]
        for fn in ("rt_part_rows", "rt_part_bytes"):
            getattr(L, fn).restype = C.c_int64
            getattr(L, fn).argtypes = [C.POINTER(RtCamera), C.c_int, C.c_int, C.c_int]
        for fn in ("rt_render_part", "rt_render_part_into_frame"):
            getattr(L, fn).argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.POINTER(RtStats)]
        L.rt_assemble_parts.argtypes = [C.POINTER(RtCamera), C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.rt_render_part_to_host.argtypes = [C.c_void_p, C.POINTER(RtCamera), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(RtStats)]
        L.rt_host_frame_create.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.rt_host_frame_open.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.rt_host_frame_close.argtypes = [C.c_void_p, C.c_int64, C.c_char_p]
        L.rt_render_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(RtCamera), C.c_int, C.c_void_p, C.POINTER(RtStats)]
        L.rt_selftest_div3.argtypes = [C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.rt_device_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p)]
        L.rt_device_free.argtypes = [C.c_void_p]
        L.rt_ipc_export.argtypes = [C.c_void_p, C.c_char_p]
        L.rt_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.rt_ipc_close.argtypes = [C.c_void_p]
        _cuda = L
    return _cuda


def _check(rc):
    if rc != 0:
        raise RtError(f"rt_b200 error {rc}: {cuda_lib().rt_last_error().decode()}")


def _host_ptr(out):
    return out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()


class RayTracer:
    """RayTracer(scene) / render(camera) — raytracer.cpp:335, :362 — on the current CUDA device."""

    def __init__(self, scene, builder=RT_BUILD_DEFAULT, brute_force=False, exact_culling=True, refill_threshold=0,
                 ploc_radius=0, ploc_leaf_cost=0.0, force_replay=False, max_ctas_per_sm=0, reinsert_rounds=0, reinsert_accept=0.0):
        self.L = cuda_lib()
        self.scene = scene
        opts = RtBuildOptions(builder, 1 if brute_force else 0, 0 if exact_culling else 1, refill_threshold, ploc_radius, ploc_leaf_cost,
                              1 if force_replay else 0, max_ctas_per_sm, reinsert_rounds, reinsert_accept)
        h = C.c_void_p()
        _check(self.L.rt_scene_create(C.byref(scene.desc), C.byref(opts), C.byref(h)))
        self.h = h
        self.last_stats = None

    def close(self):
        if getattr(self, "h", None):
            self.L.rt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        inf = RtSceneInfo()
        _check(self.L.rt_scene_info(self.h, C.byref(inf)))
        return inf

    def render(self, camera, aa=1, out=None):
        """Full frame into host memory: HxWx3 uint8 (numpy, or a pinned torch tensor passed as `out`)."""
        if out is None:
            out = np.empty((camera.image_height, camera.image_width, 3), np.uint8)
        st = RtStats()
        _check(self.L.rt_render(self.h, C.byref(camera), aa, _host_ptr(out), C.byref(st)))
        self.last_stats = st
        return out

    def render_async(self, camera, aa, out):
        """Enqueue a frame (rt_render_async); returns the ticket for wait()."""
        t = C.c_int()
        _check(self.L.rt_render_async(self.h, C.byref(camera), aa, _host_ptr(out), C.byref(t)))
        return t.value

    def wait(self, ticket):
        st = RtStats()
        _check(self.L.rt_wait(self.h, ticket, C.byref(st)))
        self.last_stats = st
        return st

    def part_bytes(self, camera, aa, rank, world):
        return int(self.L.rt_part_bytes(C.byref(camera), aa, rank, world))

    def render_part(self, camera, aa, rank, world, d_rows_ptr, stream=0, want_stats=True):
        """This GPU's interleaved row bands, packed, into DEVICE memory (e.g. a torch uint8 tensor's data_ptr())."""
        st = RtStats()
        _check(self.L.rt_render_part(self.h, C.byref(camera), aa, rank, world, d_rows_ptr, stream,
                                     C.byref(st) if want_stats else None))
        self.last_stats = st if want_stats else None
        return st

    def render_part_into_frame(self, camera, aa, rank, world, d_frame_ptr, stream=0, want_stats=True):
        st = RtStats()
        _check(self.L.rt_render_part_into_frame(self.h, C.byref(camera), aa, rank, world, d_frame_ptr, stream,
                                                C.byref(st) if want_stats else None))
        self.last_stats = st if want_stats else None
        return st

    def render_part_to_host(self, camera, aa, rank, world, host_frame_ptr):
        """This GPU's bands rendered and copied straight into their rows of a host frame (rt_render_part_to_host)."""
        st = RtStats()
        _check(self.L.rt_render_part_to_host(self.h, C.byref(camera), aa, rank, world, host_frame_ptr, C.byref(st)))
        self.last_stats = st
        return st

    def assemble(self, camera, aa, world, d_parts_ptr, part_stride, d_frame_ptr, stream=0):
        _check(self.L.rt_assemble_parts(C.byref(camera), aa, world, d_parts_ptr, part_stride, d_frame_ptr, stream))


# ----------------------------------------------------------------------------- multi-GPU plumbing (one process per GPU)
#
# The frame is cut into bands of band_height(camera, aa, world) pixel rows; band b belongs to rank b % world (the
# reference deals rows round-robin to its threads the same way, raytracer.cpp:353).  Every rank renders its bands
# into a packed buffer [n_my_bands][band_h][width][3].  For a frame that has to end up on GPU 0, rank 0 gathers the
# buffers (one collective per frame, no other data-path communication) and scatters them into the row-major
# frame; for a frame that has to end up on the HOST, no gather is needed at all: every rank copies its own bands
# over its own PCIe link into their rows of a shared page-locked frame (SharedHostFrame, rt_render_part_to_host).


def band_height(camera, aa, world):
    return int(cuda_lib().rt_band_height(C.byref(camera), aa, world))


def part_band_ids(height, band_h, rank, world):
    return range(rank, (height + band_h - 1) // band_h, world)


def gather_parts(dist, my_rows, all_parts, rank, dst=0):
    """One gather of the packed band buffers to `dst` (torch.distributed; NCCL on GPUs, gloo in the CPU tests).
    all_parts is a [world, stride] tensor on dst, None elsewhere."""
    dist.gather(my_rows, list(all_parts.unbind(0)) if rank == dst else None, dst=dst)


def pack_bands_host(frame, band_h, rank, world):
    """Host restatement of the packed layout rt_render_part writes (tests and debugging only)."""
    h, w, _ = frame.shape
    ids = part_band_ids(h, band_h, rank, world)
    out = np.zeros((len(ids), band_h, w, 3), np.uint8)
    for i, b in enumerate(ids):
        blk = frame[b * band_h:(b + 1) * band_h]
        out[i, :blk.shape[0]] = blk
    return out.reshape(-1)


def assemble_bands_host(parts, width, height, band_h, world):
    """Host restatement of rt_assemble_parts: parts is [world, stride] uint8."""
    frame = np.zeros((height, width, 3), np.uint8)
    for r in range(world):
        ids = part_band_ids(height, band_h, r, world)
        rows = np.asarray(parts[r][:len(ids) * band_h * width * 3]).reshape(len(ids), band_h, width, 3)
        for i, b in enumerate(ids):
            hh = min(band_h, height - b * band_h)
            frame[b * band_h:b * band_h + hh] = rows[i, :hh]
    return frame


class SharedHostFrame:
    """A row-major RGB8 frame in POSIX shared memory, page-locked in every rank (rt_host_frame_*): each rank's
    rt_render_part_to_host copies its bands straight into it, the frame is complete after a barrier."""

    def __init__(self, dist, rank, nbytes, root=0):
        self.L = cuda_lib()
        self.rank, self.root, self.nbytes = rank, root, nbytes
        self.ptr = C.c_void_p()
        payload = [None]
        if rank == root:
            self.name = f"/rtb200_frame_{os.getpid()}"
            _check(self.L.rt_host_frame_create(self.name.encode(), nbytes, C.byref(self.ptr)))
            payload = [self.name]
        if dist is not None:
            dist.broadcast_object_list(payload, src=root)
        self.name = payload[0]
        if rank != root:
            _check(self.L.rt_host_frame_open(self.name.encode(), nbytes, C.byref(self.ptr)))

    def as_numpy(self):
        buf = (C.c_ubyte * self.nbytes).from_address(self.ptr.value)
        return np.frombuffer(buf, dtype=np.uint8)

    def close(self):
        if self.ptr:
            self.L.rt_host_frame_close(self.ptr, self.nbytes, self.name.encode() if self.rank == self.root else None)
            self.ptr = C.c_void_p()


class PeerFrame:
    """The row-major RGB8 frame on the gathering GPU, mapped into every rank (CUDA IPC over NVLink) so that
    rt_render_part_into_frame stores finished pixels straight into it: the gather is fused into the kernel."""

    def __init__(self, dist, rank, nbytes, root=0):
        self.L = cuda_lib()
        self.rank, self.root, self.nbytes = rank, root, nbytes
        self.ptr = C.c_void_p()
        payload = [None]
        if rank == root:
            _check(self.L.rt_device_alloc(nbytes, C.byref(self.ptr)))
            h = C.create_string_buffer(64)
            _check(self.L.rt_ipc_export(self.ptr, h))
            payload = [h.raw]
        dist.broadcast_object_list(payload, src=root)
        if rank != root:
            _check(self.L.rt_ipc_open(payload[0], C.byref(self.ptr)))

    def as_tensor(self):
        """torch uint8 view of the frame (root only)."""
        import torch

        class _Ext:
            pass
        e = _Ext()
        e.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr.value, False), "version": 2}
        return torch.as_tensor(e, device="cuda")

    def close(self):
        if self.ptr:
            if self.rank == self.root:
                self.L.rt_device_free(self.ptr)
            else:
                self.L.rt_ipc_close(self.ptr)
            self.ptr = C.c_void_p()
