#!/usr/bin/env python
"""Item size of the shared-accumulator mode (AA factors that are not multiples of 8): render-kernel ms for the default
item and for smaller ones (rt_set_partition's strip width caps the item side).  python tools/item_size_experiment.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402

L = H.rt_b200.cuda_lib()
for case in ("marbles:2048:2048:4", "horse_and_mug:2880:1440:2", "horse_and_mug:3840:1920:4", "dragon_lowres:1600:1600:2", "car:2048:1536:3", "horse_and_mug:7680:3840:1",
             "horse_and_mug:1440:720:1"):
    name, w, h, aa = case.split(":")
    sc = H.golden_scene(name)
    cam = sc.camera(0, int(w), int(h))
    rt = H.RayTracer(sc)
    row = []
    for cap in (0, 16, 8, 4, 2, 1):
        L.rt_set_partition(0, cap)
        best = 1e30
        for _ in range(4):
            rt.render(cam, int(aa))
            best = min(best, rt.last_stats.ms_render)
        row.append(f"cap {cap or 'default':>7}: {best:8.3f} ms")
    print(f"{case:28s} " + "  ".join(row), flush=True)
    rt.close()
