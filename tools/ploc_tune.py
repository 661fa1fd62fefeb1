#!/usr/bin/env python
"""PLOC parameter sweep (search radius, SAH leaf cost) against the host SAH builder."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H  # noqa: E402

cases = ["marbles:2048:2048:4", "horse_and_mug:3840:1920:4", "car:2048:1536:4", "bunny:1024:1024:4", "dragon_lowres:1600:1600:2", "low_poly:2048:2048:2"]
settings = [("sah_host", 2, None, None)] + [(f"ploc r{r} c{c}", 3, r, c) for r in (8, 16, 32, 64) for c in (1.6,)] + \
           [(f"ploc r16 c{c}", 3, 16, c) for c in (1.0, 2.5, 1e10)]
for case in cases:
    name, w, h, aa = case.split(":")
    sc = H.golden_scene(name)
    cam = sc.camera(0, int(w), int(h))
    for label, b, r, c in settings:
        rt = H.RayTracer(sc, builder=b, ploc_radius=r or 0, ploc_leaf_cost=c or 0.0)
        best = 1e30
        for _ in range(4):
            rt.render(cam, int(aa))
            best = min(best, rt.last_stats.ms_render)
        inf = rt.info()
        print(f"{case:30s} {label:16s} {best:8.3f} ms  sah {inf.bvh_sah_cost:6.2f} depth {inf.bvh_max_depth:2d} nodes {inf.bvh_nodes:6d} build {inf.ms_build_device:6.2f} ms", flush=True)
        rt.close()
