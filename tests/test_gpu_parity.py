"""GPU parity: the CUDA path, called through the C-ABI (rt_b200.RayTracer -> rt_render), against the golden
frames rendered by the unmodified reference, and against the oracle on seeded/ad-hoc inputs.

Tolerance (BASELINE.json north_star): |delta| <= 1 per 8-bit channel on >= 99.9 % of pixels and zero pixels off by
more than 8.  The design goal is stricter — byte identity — and each test prints the exact counts.
"""
import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu

ALL_IMAGES = sorted(k for k, v in H.manifest()["images"].items() if "rows" not in v and not v.get("slow"))

_tracers = {}


def tracer(scene_name, **kw):
    key = (scene_name, tuple(sorted(kw.items())))
    if key not in _tracers:
        _tracers[key] = H.RayTracer(H.golden_scene(scene_name), **kw)
    return _tracers[key]


@pytest.mark.parametrize("key", ALL_IMAGES)
def test_golden_frame(key):
    gold, m = H.golden_image(key)
    sc = H.golden_scene(m["scene"])
    cam = sc.camera(m["camera"], m["width"], m["height"])
    rt = tracer(m["scene"])
    img = rt.render(cam, m["aa"])
    rep = H.diff_report(gold, img)
    st = rt.last_stats
    print(key, rep, "rays", st.primary_rays, st.reflection_rays, st.shadow_rays, st.shadow_occluded, f"{st.ms_render:.3f} ms")
    assert H.within_tolerance(rep), rep
    # Device ray counters are exact known answers: with the reference-visibility check every
    # hit decision, including the reference's own box-culling "seam holes", is reproduced.
    rays = m["rays"]
    print("   replayed rays:", st.replayed_closest, "closest,", st.replayed_any, "any-hit of", st.total_rays)
    assert (st.primary_rays, st.reflection_rays, st.shadow_rays, st.shadow_occluded) == \
        (rays["primary"], rays["reflection"], rays["shadow"], rays["shadow_occluded"])
    assert rep["equal"] == rep["pixels"]


@pytest.mark.parametrize("key", [k for k in ALL_IMAGES if k.endswith(".aa1")])
def test_bvh_equals_brute_force(key):
    """A conservative BVH must report exactly the hit set of testing every primitive (the padded boxes never cull
    what the exact tests accept): every shipped camera at its native resolution, BVH against brute force with the
    reference-visibility check switched off on both sides (the fast traversal alone).  With the check switched on,
    brute force must reproduce the reference's frame: the replay logic does not depend on the traversal tree."""
    gold, m = H.golden_image(key)
    sc = H.golden_scene(m["scene"])
    cam = sc.camera(m["camera"])
    a = tracer(m["scene"], exact_culling=False).render(cam, 1)
    b = tracer(m["scene"], brute_force=True, exact_culling=False).render(cam, 1)
    assert np.array_equal(a, b)
    c = tracer(m["scene"], brute_force=True).render(cam, 1)
    assert np.array_equal(c, gold)


def test_tiles_equal_full_frame():
    """Interleaved-tile parts (packed and straight-into-frame) reassemble to the single-GPU frame byte for byte."""
    import torch
    sc = H.golden_scene("cornellbox")
    cam = sc.camera(0, 250, 190)  # not a multiple of the tile size
    rt = tracer("cornellbox")
    full = rt.render(cam, 2)
    for world in (2, 3, 8):
        stride = rt.part_bytes(cam, 0, world)
        parts = torch.zeros(world * stride, dtype=torch.uint8, device="cuda")
        frame = torch.zeros(cam.image_height * cam.image_width * 3, dtype=torch.uint8, device="cuda")
        frame2 = torch.zeros_like(frame)
        total = 0
        for r in range(world):
            st = rt.render_part(cam, 2, r, world, parts.data_ptr() + r * stride)
            total += st.primary_rays
            rt.render_part_into_frame(cam, 2, r, world, frame2.data_ptr())
        rt.assemble(cam, world, parts.data_ptr(), stride, frame.data_ptr())
        torch.cuda.synchronize()
        got = frame.cpu().numpy().reshape(full.shape)
        got2 = frame2.cpu().numpy().reshape(full.shape)
        assert np.array_equal(got, full), world
        assert np.array_equal(got2, full), world
        assert total == cam.image_width * cam.image_height * 4


def test_oracle_on_odd_configuration():
    """Ad-hoc camera/AA the goldens do not cover, against the oracle rendered on the box's CPU."""
    sc = H.golden_scene("simple_reflectance")
    cam = sc.camera(0, 97, 61)
    orc = H.OracleScene(sc)
    want, ost = orc.render(cam, 7)
    rt = tracer("simple_reflectance")
    got = rt.render(cam, 7)
    rep = H.diff_report(want, got)
    print(rep)
    assert H.within_tolerance(rep)
    assert rt.last_stats.primary_rays == ost.primary_rays


@pytest.mark.parametrize("key", ["simple.aa1", "cornellbox_front.aa1", "marbles.aa1", "bunny.aa1", "horse_and_mug.aa1",
                                 "dragon_lowres.aa1", "low_poly_scene.aa1", "mirror_spheres.aa1", "Car.aa1"])
@pytest.mark.parametrize("builder", ["lbvh", "ploc", "sah_gpu"])
def test_gpu_bvh_builders(key, builder):
    """The BVHs built on the GPU (Morton + bitonic sort, then Karras' radix tree or PLOC clustering, SAH leaf collapse)
    must give the same frames: the tree is only a filter, ties are settled by the reference-order ranks."""
    gold, m = H.golden_image(key)
    sc = H.golden_scene(m["scene"])
    rt = tracer(m["scene"], builder={"lbvh": H.rt_b200.RT_BUILD_LBVH_GPU, "ploc": H.rt_b200.RT_BUILD_PLOC_GPU, "sah_gpu": H.rt_b200.RT_BUILD_SAH_GPU}[builder])
    img = rt.render(sc.camera(m["camera"]), 1)
    inf, ref_inf = rt.info(), tracer(m["scene"], builder=H.rt_b200.RT_BUILD_SAH_HOST).info()
    print(key, H.diff_report(gold, img), f"{builder}: {inf.bvh_nodes} nodes depth {inf.bvh_max_depth} sah {inf.bvh_sah_cost:.1f} "
          f"build {inf.ms_build_device:.3f} ms device, {rt.last_stats.ms_render:.3f} ms render | sah_host: {ref_inf.bvh_nodes} nodes "
          f"sah {ref_inf.bvh_sah_cost:.1f} build {ref_inf.ms_build_host:.1f} ms host")
    assert np.array_equal(img, gold)


@pytest.mark.slow
def test_full_size_config5(tmp_path):
    """BASELINE.json config 5 at full size: horse_and_mug 7680x3840, 16x16 SSAA (25.99 G rays, ~1 s on a B200).
    Known answers: the reference's ray counts and the md5 of its PPM (64-bit-index build), and
    12 output rows rebuilt from the reference's own sub-samples (tests/golden/make_golden.py)."""
    import hashlib
    gold, m = H.golden_image("horse_and_mug_8k.aa16.rows")
    sc = H.golden_scene("horse_and_mug")
    cam = sc.camera(0, m["width"], m["height"])
    rt = tracer("horse_and_mug")
    img = rt.render(cam, m["aa"])
    st = rt.last_stats
    rays = m["rays"]
    print("8K 16x:", st.primary_rays, st.reflection_rays, st.shadow_rays, f"{st.ms_render:.1f} ms render, {st.ms_d2h:.2f} ms D2H,",
          f"{st.total_rays / st.ms_render / 1e3:.0f} Mrays/s")
    print("   replayed rays:", st.replayed_closest, "closest,", st.replayed_any, "any-hit")
    full, fm = H.golden_image("horse_and_mug_8k.aa16.full")
    bad = np.argwhere((full != img).any(axis=2))
    print("   pixels differing from the reference frame:", len(bad), [(int(y), int(x), full[y, x].tolist(), img[y, x].tolist()) for y, x in bad[:20]])
    assert st.primary_rays == rays["primary"] == 7680 * 3840 * 256
    assert (st.reflection_rays, st.shadow_rays) == (rays["reflection"], rays["shadow"])
    got = img[m["rows"]]
    rep = H.diff_report(gold, got)
    print("golden rows:", rep)
    assert H.within_tolerance(rep)
    # the whole frame against the committed full-size golden (oracle frame whose PPM md5 equals the reference's)
    full, fm = H.golden_image("horse_and_mug_8k.aa16.full")
    frep = H.diff_report(full, img)
    print("full frame vs reference:", frep)
    assert H.within_tolerance(frep)
    assert frep["equal"] == frep["pixels"]
    p = str(tmp_path / "h8k.ppm")
    H.write_ppm(p, img)
    h = hashlib.md5()
    with open(p, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    identical = h.hexdigest() == fm["ppm_md5"]
    print("full-frame PPM md5", h.hexdigest(), "== reference's" if identical else "!= reference's " + fm["ppm_md5"])
    assert identical == (frep["equal"] == frep["pixels"])
    img2 = rt.render(cam, m["aa"])
    assert np.array_equal(img, img2)


def test_cli_matches_reference_binary(tmp_path):
    """`raytracer scene.xml` (default 2x2 SSAA, as the reference ships) writes the same PPM bytes as the reference's
    own binary built by its Makefile flags (oracle/_ref/raytracer), and prints the reference's timing lines."""
    import os
    import subprocess
    ours = os.path.join(H.PKG, "raytracer")
    ref = os.path.join(H.ROOT, "oracle", "_ref", "raytracer")
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("CLI binaries not built")
    for scene in ("simple_reflectance", "cornellbox"):  # one and three cameras
        xml = H.golden_scene_path(scene)
        a, b = tmp_path / (scene + "_ours"), tmp_path / (scene + "_ref")
        a.mkdir()
        b.mkdir()
        out = subprocess.run([ours, xml, "--stats"], cwd=a, capture_output=True, text=True, check=True).stdout
        subprocess.run([ref, xml], cwd=b, capture_output=True, text=True, check=True)
        assert "Planted trees in" in out and "Rendered in" in out and "Total:" in out
        assert "Super Sampling Anti aliasing is enabled. (2*2x)" in out
        names = sorted(os.listdir(b))
        assert names and sorted(os.listdir(a)) == names
        for n in names:
            assert open(a / n, "rb").read() == open(b / n, "rb").read(), (scene, n)


def test_single_process_multi_gpu():
    """rt_render_multi (what `raytracer --gpus N` uses): one handle per device, interleaved tiles, peer copies to
    device 0, one D2H — the frame must equal the single-GPU frame.  Needs >= 2 GPUs (gpurun --gpus 2)."""
    import ctypes as C
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    L = H.rt_b200.cuda_lib()
    sc = H.golden_scene("horse_and_mug")
    cam = sc.camera(0, 700, 390)
    single = tracer("horse_and_mug").render(cam, 3)
    want = tracer("horse_and_mug").last_stats
    handles = []
    for d in range(n):
        assert L.rt_set_device(d) == 0
        handles.append(H.RayTracer(sc))
    L.rt_set_device(0)
    arr = (C.c_void_p * n)(*[h.h for h in handles])
    out = np.zeros_like(single)
    st = H.RtStats()
    rc = L.rt_render_multi(arr, n, C.byref(cam), 3, out.ctypes.data, C.byref(st))
    assert rc == 0, L.rt_last_error()
    assert np.array_equal(out, single)
    assert (st.primary_rays, st.reflection_rays, st.shadow_rays) == (want.primary_rays, want.reflection_rays, want.shadow_rays)
    for h in handles:
        h.close()


@pytest.mark.parametrize("seed", list(range(1, 13)))
def test_seeded_scenes_against_oracle(seed):
    """Seeded triangle/sphere soups with the hard cases mixed in (zero-thickness boxes, exact-t ties on shared
    edges, a degenerate triangle, mirrors, the camera inside a sphere on even seeds), all three kernels and all
    three BVH builders against the oracle on the box's CPU: byte identity and equal ray counts."""
    sc = H.random_scene(seed, n_tris=30 + 7 * seed, camera_inside_sphere=(seed % 2 == 0), max_depth=seed % 5, width=120, height=72)
    cam = sc.camera(0)
    aa = 1 + seed % 3
    want, ost = H.OracleScene(sc).render(cam, aa)
    import os
    builders = (H.rt_b200.RT_BUILD_PLOC_GPU, H.rt_b200.RT_BUILD_SAH_GPU, H.rt_b200.RT_BUILD_SAH_HOST) if seed % 2 else \
        (H.rt_b200.RT_BUILD_LBVH_GPU, H.rt_b200.RT_BUILD_SAH_GPU, H.rt_b200.RT_BUILD_AUTO)
    for kernel, builder in (("2", builders[seed % 3]), ("1", builders[(seed + 1) % 3]), ("3", builders[(seed + 2) % 3])):
        os.environ["RT_B200_KERNEL"] = kernel
        try:
            rt = H.RayTracer(sc, builder=builder)
        finally:
            os.environ.pop("RT_B200_KERNEL", None)
        got = rt.render(cam, aa)
        st = rt.last_stats
        rep = H.diff_report(want, got)
        assert rep["equal"] == rep["pixels"], (seed, kernel, builder, rep)
        assert (st.primary_rays, st.reflection_rays, st.shadow_rays, st.shadow_occluded) == \
            (ost.primary_rays, ost.reflection_rays, ost.shadow_rays, ost.shadow_occluded), (seed, kernel, builder)
        rt.close()


def _tiny_scene(tris, spheres=(), lights=((0, 5, 0, 500, 500, 500),), depth=2, mirror=1):
    verts = [[-1, -1, -5], [1, -1, -5], [0, 1, -5], [0, 0, -3]]
    m13 = [[0.2, 0.2, 0.2, 0.5, 0.5, 0.5, 0.3, 0.3, 0.3, 0.8, 0.8, 0.8, 10]]
    cam = H.RtCamera(H.RtVec3(0, 0, 0), H.RtVec3(0, 0, -1), H.RtVec3(0, 1, 0), -1, 1, -1, 1, 1, 33, 17)
    return H.Scene(np.array(verts, np.float32), np.array(tris, np.int32).reshape(-1, 4), np.array([s[:2] for s in spheres], np.int32).reshape(-1, 2),
                   np.array([s[2] for s in spheres], np.float32), m13, [mirror], np.array(lights, np.float32).reshape(-1, 6), [10, 10, 10], 1e-3,
                   [7, 8, 9], depth, [(cam, "tiny.ppm")])


@pytest.mark.parametrize("case", ["empty", "no_lights", "one_triangle", "one_sphere", "depth0", "depth32", "one_pixel", "aa64"])
def test_edge_cases_against_oracle(case):
    """Degenerate inputs the reference handles implicitly: no primitives (background only), no lights (ambient
    only), single-primitive trees, recursion depth 0 and the supported maximum, a 1x1 frame, a 64x64 sample grid."""
    sc = {"empty": lambda: _tiny_scene([]), "no_lights": lambda: _tiny_scene([[1, 2, 3, 1]], lights=()),
          "one_triangle": lambda: _tiny_scene([[1, 2, 3, 1]]), "one_sphere": lambda: _tiny_scene([], spheres=[(1, 4, 0.7)]),
          "depth0": lambda: _tiny_scene([[1, 2, 3, 1]], spheres=[(1, 4, 0.7)], depth=0),
          "depth32": lambda: _tiny_scene([[1, 2, 3, 1]], spheres=[(1, 4, 0.7)], depth=32),
          "one_pixel": lambda: _tiny_scene([[1, 2, 3, 1]]), "aa64": lambda: _tiny_scene([[1, 2, 3, 1]], spheres=[(1, 4, 0.7)])}[case]()
    cam = sc.camera(0, 1, 1) if case == "one_pixel" else (sc.camera(0, 5, 3) if case == "aa64" else sc.camera(0))
    aa = 64 if case == "aa64" else 2
    want, ost = H.OracleScene(sc).render(cam, aa)
    rt = H.RayTracer(sc)
    got = rt.render(cam, aa)
    assert np.array_equal(want, got), H.diff_report(want, got)
    assert rt.last_stats.total_rays == ost.total_rays
    rt.close()


def test_argument_errors():
    """The C-ABI returns error codes instead of crashing: bad ids, bad AA factor, unsupported recursion depth."""
    sc = _tiny_scene([[1, 2, 9, 1]])  # vertex id 9 does not exist
    with pytest.raises(H.rt_b200.RtError, match="vertex id out of range"):
        H.RayTracer(sc)
    with pytest.raises(H.rt_b200.RtError, match="max_recursion_depth"):
        H.RayTracer(_tiny_scene([[1, 2, 3, 1]], depth=33))
    rt = H.RayTracer(_tiny_scene([[1, 2, 3, 1]]))
    with pytest.raises(H.rt_b200.RtError, match="aa_factor"):
        rt.render(rt.scene.camera(0), 0)
    rt.close()


@pytest.mark.parametrize("scene", sorted(H.manifest()["scenes"]))
def test_reference_tree_on_gpu_equals_host(scene):
    """The reference-order tree and the 8 rank arrays built on the GPU (ref_order_device.cu, what rt_scene_create
    uses) against the host build (ref_order.cpp), whose statistics equal the reference's own builder's: ranks, node /
    leaf / depth statistics and an FNV hash over every node box, link and leaf list must be identical."""
    import ctypes as C
    L = H.rt_b200.cuda_lib()
    L.rt_host_reference_ranks.argtypes = [C.POINTER(H.RtSceneDesc), C.c_void_p, C.c_void_p]
    L.rt_host_reference_tree_hash.argtypes = [C.POINTER(H.RtSceneDesc), C.POINTER(C.c_uint64)]
    L.rt_device_reference_ranks.argtypes = [C.POINTER(H.RtSceneDesc), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    sc = H.golden_scene(scene)
    n = sc.desc.n_triangles + sc.desc.n_spheres
    hr, dr = np.zeros((8, n), np.uint32), np.zeros((8, n), np.uint32)
    hs, ds = (C.c_int32 * 4)(), (C.c_int32 * 4)()
    hh, dh = C.c_uint64(), C.c_uint64()
    assert L.rt_host_reference_ranks(C.byref(sc.desc), hr.ctypes.data, hs) == 0
    assert L.rt_host_reference_tree_hash(C.byref(sc.desc), C.byref(hh)) == 0
    assert L.rt_device_reference_ranks(C.byref(sc.desc), dr.ctypes.data, ds, C.byref(dh)) == 0, L.rt_last_error()
    assert list(hs) == list(ds)
    assert dict(zip(["nodes", "leaves", "max_leaf", "max_depth"], list(ds))) == H.manifest()["scenes"][scene]["ref_bvh"]
    assert np.array_equal(hr, dr)
    assert hh.value == dh.value


@pytest.mark.slow
@pytest.mark.parametrize("scene,w,h,aa", [("dragon_lowres", 2400, 2400, 3), ("mirror_spheres", 3072, 3072, 2), ("marbles", 1024, 1024, 3),
                                          ("car", 2048, 1536, 2), ("berserker", 1536, 2048, 2)])
def test_large_frames_against_oracle(scene, w, h, aa):
    """BASELINE.json config 4 ("high-triangle and deep-recursion path") and friends at sizes nobody pre-rendered:
    the GPU frame against the oracle rendered on the box's host cores (a few seconds each) — byte identity and
    equal ray counts, tens to hundreds of millions of rays per case."""
    sc = H.golden_scene(scene)
    cam = sc.camera(0, w, h)
    want, ost = H.OracleScene(sc).render(cam, aa)
    rt = tracer(scene)
    got = rt.render(cam, aa)
    st = rt.last_stats
    rep = H.diff_report(want, got)
    print(scene, rep, "rays", st.total_rays, "replayed", st.replayed_closest, st.replayed_any, f"{st.ms_render:.2f} ms")
    assert rep["equal"] == rep["pixels"], rep
    assert (st.primary_rays, st.reflection_rays, st.shadow_rays, st.shadow_occluded) == \
        (ost.primary_rays, ost.reflection_rays, ost.shadow_rays, ost.shadow_occluded)


@pytest.mark.parametrize("scene,cam_idx,w,h,aa", [("cornellbox", 0, 101, 101, 1), ("cornellbox", 2, 51, 51, 3), ("simple", 0, 33, 33, 1),
                                                  ("mirror_spheres", 0, 65, 65, 1), ("simple_reflectance", 0, 9, 9, 5)])
def test_axis_parallel_rays(scene, cam_idx, w, h, aa):
    """Odd resolutions on symmetric near planes put pixel centres exactly on the optical axis: direction components
    that are exactly 0 (1/d = inf), rays inside the planes of zero-thickness boxes.  The fast box test clamps the
    reciprocal, the replay uses the reference's inf/NaN arithmetic — both must agree with the oracle."""
    sc = H.golden_scene(scene)
    cam = sc.camera(cam_idx, w, h)
    want, ost = H.OracleScene(sc).render(cam, aa)
    rt = tracer(scene)
    got = rt.render(cam, aa)
    st = rt.last_stats
    assert np.array_equal(want, got), H.diff_report(want, got)
    assert (st.primary_rays, st.reflection_rays, st.shadow_rays, st.shadow_occluded) == \
        (ost.primary_rays, ost.reflection_rays, ost.shadow_rays, ost.shadow_occluded)
