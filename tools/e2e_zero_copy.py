#!/usr/bin/env python
"""End-to-end rt_render into page-locked memory: kernel stores straight into the host frame (default) against render into
device memory + D2H copy (rt_set_zero_copy(-1)).  Median of 30 calls (3 for the 8K frame)."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

L = H.rt_b200.cuda_lib()
L.rt_set_zero_copy.argtypes = [__import__("ctypes").c_int64]
cases = [("simple", 800, 800, 1), ("bunny", 512, 512, 1), ("horse_and_mug", 1440, 720, 1), ("dragon_lowres", 800, 800, 1), ("mirror_spheres", 1024, 1024, 1),
         ("horse_and_mug", 1440, 720, 2), ("horse_and_mug", 1920, 960, 16), ("horse_and_mug", 7680, 3840, 16)]
for name, w, h, aa in cases:
    sc = H.golden_scene(name)
    cam = sc.camera(0, w, h)
    rt = H.RayTracer(sc)
    out = torch.empty(w * h * 3, dtype=torch.uint8, pin_memory=True)
    res, frames = {}, {}
    for mode, limit in (("copy", -1), ("zero-copy", 0), ("copy", -1), ("zero-copy", 0)):
        L.rt_set_zero_copy(limit)
        reps = 3 if w >= 7680 else 30
        for _ in range(2):
            rt.render(cam, aa, out=out)
        t = []
        k = []
        for _ in range(reps):
            t0 = time.perf_counter()
            rt.render(cam, aa, out=out)
            t.append((time.perf_counter() - t0) * 1e3)
            k.append(rt.last_stats.ms_render)
        res.setdefault(mode, []).append((statistics.median(t), statistics.median(k)))
        frames[mode] = out.numpy().copy()
    same = bool(np.array_equal(frames["copy"], frames["zero-copy"]))
    print(f"{name:15s} {w}x{h} aa{aa}: copy e2e {min(x[0] for x in res['copy']):8.3f} ms (kernel {min(x[1] for x in res['copy']):8.3f})   "
          f"zero-copy e2e {min(x[0] for x in res['zero-copy']):8.3f} ms (kernel {min(x[1] for x in res['zero-copy']):8.3f})   identical={same}", flush=True)
    rt.close()
