#!/usr/bin/env python
"""A/B of the BVH builders and the exact-culling switch.  Usage: python tools/builder_ab.py [scene:width:height:aa ...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H  # noqa: E402

cases = sys.argv[1:] or ["marbles:2048:2048:4", "horse_and_mug:3840:1920:4", "car:2048:1536:4", "bunny:1024:1024:4", "dragon_lowres:1600:1600:2"]
for case in cases:
    name, w, h, aa = case.split(":")
    sc = H.golden_scene(name)
    cam = sc.camera(0, int(w), int(h))
    for bname, b in (("ploc", 3), ("sah_gpu", 5), ("sah_host", 2), ("lbvh", 1)):
        for exact in (True, False):
            rt = H.RayTracer(sc, builder=b, exact_culling=exact)
            best = 1e30
            for _ in range(4):
                rt.render(cam, int(aa))
                best = min(best, rt.last_stats.ms_render)
            st, inf = rt.last_stats, rt.info()
            print(f"{case:30s} {bname:8s} exact={exact!s:5s} {best:8.3f} ms {st.total_rays / best / 1e3:8.0f} Mrays/s  sah {inf.bvh_sah_cost:6.1f} "
                  f"depth {inf.bvh_max_depth:2d} replays {st.replayed_closest}+{st.replayed_any}", flush=True)
            rt.close()
