#!/usr/bin/env python
"""Where scene creation and the first frame spend their time, cold and warm (no torch in the process: the library's
own CUDA runtime brings the context up).  One JSON line per scene.

    python tools/scene_create_profile.py [scene ...]
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import numpy as np  # noqa: E402


def main():
    scenes = sys.argv[1:] or ["horse_and_mug", "simple", "bunny", "dragon_lowres", "mirror_spheres", "marbles"]
    L = H.rt_b200.cuda_lib()
    t0 = time.perf_counter()
    n = L.rt_device_count()
    L.rt_set_device(0)
    t_ctx = time.perf_counter() - t0
    print(json.dumps({"devices": n, "device_count_plus_set_device_s": t_ctx, "module_loading": os.environ.get("CUDA_MODULE_LOADING", "default")}))
    for name in scenes:
        sc = H.golden_scene(name)
        cam = sc.camera(0)
        out = np.empty((cam.image_height, cam.image_width, 3), np.uint8)
        rec = {"scene": name, "create_s": [], "ms_build_host": [], "ms_build_device": [], "render_s": [], "ms_render": [], "ms_d2h": []}
        for i in range(4):
            t0 = time.perf_counter()
            rt = H.RayTracer(sc)
            rec["create_s"].append(round(time.perf_counter() - t0, 5))
            inf = rt.info()
            rec["ms_build_host"].append(round(inf.ms_build_host, 3))
            rec["ms_build_device"].append(round(inf.ms_build_device, 3))
            for j in range(3 if i == 0 else 2):
                t0 = time.perf_counter()
                rt.render(cam, 1, out=out)
                rec["render_s"].append(round(time.perf_counter() - t0, 5))
                rec["ms_render"].append(round(rt.last_stats.ms_render, 4))
                rec["ms_d2h"].append(round(rt.last_stats.ms_d2h, 4))
            t0 = time.perf_counter()
            rt.close()
            rec.setdefault("destroy_s", []).append(round(time.perf_counter() - t0, 5))
        rec["builder"] = inf.builder
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
