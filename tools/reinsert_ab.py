#!/usr/bin/env python
"""A/B of the insertion-based tree optimisation (bvh_reinsert.cu) on the GPU: render time, SAH cost, build time.
Usage: python tools/reinsert_ab.py [scene:width:height:aa ...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H  # noqa: E402

B = H.rt_b200
cases = sys.argv[1:] or ["horse_and_mug:3840:1920:16", "horse_and_mug:1440:720:1", "car:2048:1536:4", "bunny:1024:1024:4",
                         "dragon_lowres:1600:1600:2", "marbles:2048:2048:4", "cornellbox:1600:1600:4", "mirror_spheres:2048:2048:2"]
variants = [("auto (default)", dict(builder=B.RT_BUILD_AUTO)),
            ("auto, no reinsertion", dict(builder=B.RT_BUILD_AUTO, reinsert_rounds=-1)),
            ("ploc", dict(builder=B.RT_BUILD_PLOC_GPU)),
            ("sah plain", dict(builder=B.RT_BUILD_SAH_GPU, reinsert_rounds=-1)),
            ("sah + 4 rounds, forced", dict(builder=B.RT_BUILD_SAH_GPU, reinsert_rounds=4, reinsert_accept=1e9)),
            ("sah + 8 rounds, forced", dict(builder=B.RT_BUILD_SAH_GPU, reinsert_rounds=8, reinsert_accept=1e9)),
            ("sah + 24 rounds, forced", dict(builder=B.RT_BUILD_SAH_GPU, reinsert_rounds=24, reinsert_accept=1e9))]
for case in cases:
    name, w, h, aa = case.split(":")
    sc = H.golden_scene(name)
    cam = sc.camera(0, int(w), int(h))
    ref = None
    for vname, kw in variants:
        builds = []
        for _ in range(3):
            rt = H.RayTracer(sc, **kw)
            builds.append(rt.info().ms_build_device)
            if _ < 2:
                rt.close()
        best = 1e30
        for _ in range(5):
            img = rt.render(cam, int(aa))
            best = min(best, rt.last_stats.ms_render)
        if ref is None:
            ref = img.copy()
        st, inf = rt.last_stats, rt.info()
        print(f"{case:28s} {vname:26s} {best:9.3f} ms {st.total_rays / best / 1e3:8.0f} Mrays/s  kept {inf.builder} sah {inf.bvh_sah_cost:6.3f} depth {inf.bvh_max_depth:2d} "
              f"nodes {inf.bvh_nodes:6d} | reinsert {inf.reinsert_moves:4d} moves/{inf.reinsert_rounds:2d} rounds {inf.reinsert_cost_before:6.3f}->{inf.reinsert_cost_after:6.3f} "
              f"kept {inf.reinsert_accepted} | build {min(builds):6.3f} ms | same frame {bool((img == ref).all())}", flush=True)
        rt.close()
