#!/usr/bin/env python
"""Tiny renders for compute-sanitizer (one tool per run):  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H  # noqa: E402

for name, w, h, aa in (("simple_reflectance", 64, 48, 2), ("cornellbox", 40, 40, 3), ("bunny", 48, 48, 1), ("marbles", 32, 32, 2)):
    sc = H.golden_scene(name)
    cam = sc.camera(0, w, h)
    want, _ = H.OracleScene(sc).render(cam, aa)
    for builder in (0, 1, 3, 5, 2):
        for refill in (0, 12):
            rt = H.RayTracer(sc, builder=builder, refill_threshold=refill)
            got = rt.render(cam, aa)
            assert (got == want).all(), (name, builder, refill)
            if refill == 0:
                got8 = rt.render(cam, 8)  # register-accumulator strips
                assert got8.shape == want.shape
            rt.close()
    print(name, "ok", flush=True)
