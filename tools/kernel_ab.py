#!/usr/bin/env python
"""A/B timing of the render-kernel variants (RT_B200_KERNEL / RT_B200_REFILL) on the shipped scenes.
Usage: python tools/kernel_ab.py [scene:width:height:aa ...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H  # noqa: E402

cases = sys.argv[1:] or ["marbles:2048:2048:4", "mirror_spheres:2048:2048:4", "car:2048:1536:4", "horse_and_mug:1440:720:1",
                         "dragon_lowres:1600:1600:2", "bunny:1024:1024:4"]
variants = [("1", "0"), ("3", "0"), ("2", "0"), ("2", "8"), ("2", "16"), ("2", "24"), ("2", "31")]
for case in cases:
    name, w, h, aa = case.split(":")
    sc = H.golden_scene(name)
    cam = sc.camera(0, int(w), int(h))
    ref = None
    for k, r in variants:
        os.environ["RT_B200_KERNEL"], os.environ["RT_B200_REFILL"] = k, r
        rt = H.RayTracer(sc)
        best = 1e30
        for _ in range(4):
            img = rt.render(cam, int(aa))
            best = min(best, rt.last_stats.ms_render)
        st = rt.last_stats
        if ref is None:
            ref = img.copy()
        same = bool((img == ref).all())
        print(f"{case:32s} kernel {k} refill {r:>2s}: {best:8.3f} ms  {st.total_rays / best / 1e3:8.0f} Mrays/s  identical={same}", flush=True)
        rt.close()
