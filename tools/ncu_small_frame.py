#!/usr/bin/env python
"""One warm launch of the render kernel on a no-AA frame, for `ncu --set full -k regex:render_kernel -s N -c 1`.
    python tools/ncu_small_frame.py [scene] [aa] [width height]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "horse_and_mug"
aa = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = H.golden_scene(scene)
cam = sc.camera(0, *(int(x) for x in sys.argv[3:5])) if len(sys.argv) > 4 else sc.camera(0)
rt = H.RayTracer(sc)
for _ in range(4):
    rt.render(cam, aa)
    st = rt.last_stats
    print(scene, aa, cam.image_width, cam.image_height, "rays", st.total_rays, f"{st.ms_render:.4f} ms render, {st.ms_d2h:.4f} ms d2h")
rt.close()
