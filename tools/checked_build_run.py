#!/usr/bin/env python
"""Runs the render path and every builder of the BOUNDS-CHECKED library build (ab/checked.so: -DRT_BOUNDS_CHECK turns every
RT_CHECK into a device assert) over the shipped scenes, supersampled and partitioned cases and seeded soups, comparing with
the golden frames / the oracle.  A failed assert kills the CUDA context: this script then exits non-zero with the
assertion's file:line on stderr.  tests/test_gpu_parity.py::test_bounds_checked_build runs it in a subprocess."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("RT_B200_LIB", os.path.join(ROOT, "raytracer-ceng477-graphics-hw-1_b200", "ab", "checked.so"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

B = H.rt_b200
n = 0
for key, m in sorted(H.manifest()["images"].items()):
    if "rows" in m or m.get("slow"):
        continue
    gold, _ = H.golden_image(key)
    sc = H.golden_scene(m["scene"])
    cam = sc.camera(m["camera"], m["width"], m["height"])
    for builder in ((B.RT_BUILD_AUTO, B.RT_BUILD_PLOC_GPU, B.RT_BUILD_SAH_GPU, B.RT_BUILD_LBVH_GPU) if m["aa"] == 1 else (B.RT_BUILD_AUTO,)):
        rt = H.RayTracer(sc, builder=builder)
        img = rt.render(cam, m["aa"])
        assert np.array_equal(img, gold), (key, builder)
        rt.close()
        n += 1
sc = H.golden_scene("horse_and_mug")
rt = H.RayTracer(sc)
for aa, w, h in ((8, 333, 171), (16, 480, 240), (24, 64, 40), (3, 500, 250), (5, 123, 77)):
    cam = sc.camera(0, w, h)
    full = rt.render(cam, aa)
    frame = torch.zeros(w * h * 3, dtype=torch.uint8, device="cuda")
    host = np.zeros(w * h * 3, np.uint8)
    for world in (2, 5):
        stride = rt.part_bytes(cam, aa, 0, world)
        parts = torch.zeros(world * stride, dtype=torch.uint8, device="cuda")
        for r in range(world):
            rt.render_part(cam, aa, r, world, parts.data_ptr() + r * stride)
            rt.render_part_into_frame(cam, aa, r, world, frame.data_ptr())
            rt.render_part_to_host(cam, aa, r, world, host.ctypes.data)
        out = torch.zeros_like(frame)
        rt.assemble(cam, aa, world, parts.data_ptr(), stride, out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().reshape(full.shape), full) and np.array_equal(frame.cpu().numpy().reshape(full.shape), full)
        assert np.array_equal(host.reshape(full.shape), full)
        n += 1
rt.close()
for seed in range(300, 330):
    sc = H.random_scene(seed, n_tris=10 + 9 * (seed % 17), n_spheres=seed % 5, camera_inside_sphere=(seed % 4 == 0), max_depth=seed % 6, width=80, height=48)
    cam = sc.camera(0)
    aa = (1, 2, 8, 3)[seed % 4]
    want, _ = H.OracleScene(sc).render(cam, aa)
    for force in (False, True):
        rt = H.RayTracer(sc, builder=(B.RT_BUILD_AUTO, B.RT_BUILD_PLOC_GPU, B.RT_BUILD_LBVH_GPU)[seed % 3], force_replay=force, refill_threshold=(0 if aa == 8 else 4 * (seed % 3)))
        assert np.array_equal(rt.render(cam, aa), want), (seed, force)
        rt.close()
        n += 1
big = H.tessellated_scene("bunny", 2)  # 79 K triangles: multi-CTA PLOC
rt = H.RayTracer(big)
cam = big.camera(0, 256, 256)
want, _ = H.OracleScene(big).render(cam, 1)
assert np.array_equal(rt.render(cam, 1), want)
rt.close()
print("checked build:", n + 1, "cases, no assertion fired, every frame identical")
