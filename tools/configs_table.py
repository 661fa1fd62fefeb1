#!/usr/bin/env python
"""BASELINE.json configs 1-4 on one GPU next to the reference on the box's host cores (config 5 is bench.py).
Writes a markdown table to stdout.  Every GPU frame is compared with the committed golden frame."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H  # noqa: E402

CONFIGS = [("1", "simple", "simple.aa1"), ("2", "bunny", "bunny.aa1"), ("3", "horse_and_mug", "horse_and_mug.aa1"),
           ("4a", "dragon_lowres", "dragon_lowres.aa1"), ("4b", "mirror_spheres", "mirror_spheres.aa1"),
           ("-", "marbles", "marbles.aa1"), ("-", "car", "Car.aa1"), ("3 (as shipped, 2x2)", "horse_and_mug", "horse_and_mug.aa2")]
print(f"| config | scene (output, AA) | rays | GPU kernel ms | GPU e2e ms (frame on host) | Mrays/s (kernel) | reference render-only s ({os.cpu_count()} host cores) | CPU Mrays/s | e2e speed-up | frame vs reference |")
print("|---|---|---|---|---|---|---|---|---|---|")
for cfg, scene, key in CONFIGS:
    gold, m = H.golden_image(key)
    sc = H.golden_scene(scene)
    cam = sc.camera(m["camera"], m["width"], m["height"])
    rt = H.RayTracer(sc)
    out = np.empty_like(gold)
    best_e2e, best_k = 1e30, 1e30
    for _ in range(12):
        t0 = time.perf_counter()
        rt.render(cam, m["aa"], out=out)
        best_e2e = min(best_e2e, (time.perf_counter() - t0) * 1e3)
        best_k = min(best_k, rt.last_stats.ms_render)
    st = rt.last_stats
    same = "byte-identical" if np.array_equal(out, gold) else str(H.diff_report(gold, out))
    rt.close()
    cpu = float("nan")
    if H.ref_available():
        ref = H.RefScene(H.golden_scene_path(scene))
        cpu = min(ref.render(m["camera"], m["aa"])[1] for _ in range(3))
        ref.close()
    print(f"| {cfg} | {key.split('.')[0]} ({m['width']}x{m['height']}, {m['aa']}x{m['aa']}) | {st.total_rays} | {best_k:.3f} | {best_e2e:.3f} | "
          f"{st.total_rays / best_k / 1e3:.0f} | {cpu:.3f} | {st.total_rays / cpu / 1e6:.1f} | {cpu * 1e3 / best_e2e:.0f}x | {same} |", flush=True)
