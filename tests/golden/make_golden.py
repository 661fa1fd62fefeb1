#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference and oracle/_ref/libref.so):

    make -C oracle all && python tests/golden/make_golden.py

What it writes
  scenes/<scene>.xml.xz          the reference's shipped input scenes (inputs/*.xml), xz-compressed —
                                 they are the benchmark's named input data (BASELINE.json configs)
  images/<image>.aa<f>.rgb.xz    RGB8 frames rendered by the reference's own renderer
                                 (RayTracer::render + ImageProcessor::downSample via oracle/ref_shim.cpp)
  images/horse_and_mug_8k.aa16.rows.rgb.xz
                                 selected OUTPUT rows of the 7680x3840 16x16-SSAA frame (config 5), built
                                 from the reference's own sub-samples (ref_time_rows) and the integer
                                 average of raytracer.cpp:466-477
  images/horse_and_mug_8k.aa16.full.rgb.xz
                                 the whole config-5 frame (only with --full-8k: ~25 min of CPU).  The reference cannot
                                 render it as written (int overflow, raytracer.cpp:363), so the oracle renders it and
                                 the result is accepted only if its P3 md5 equals the md5 of the reference's 64-bit-index
                                 build recorded in SURVEY.md 8c (2be09d2f...) and its ray counts equal SURVEY.md 8d
  manifest.json                  sizes, md5 of the P3 text the reference's write_ppm emits, sha256 of the
                                 raw frames, loader digests, reference BVH statistics, and the oracle's
                                 ray counters (known answers for the GPU's device counters)
"""
import glob
import hashlib
import json
import lzma
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402

REF_INPUTS = "/root/reference/inputs"

# (scene, camera index, aa factor, width override, height override, tag)
AA_CASES = [
    ("simple", 0, 2, 0, 0, None),            # as shipped (raytracer.cpp:26-28)
    ("bunny", 0, 2, 0, 0, None),
    ("horse_and_mug", 0, 2, 0, 0, None),
    ("cornellbox", 0, 3, 0, 0, None),        # odd factor
    ("simple_reflectance", 0, 4, 0, 0, None),
    ("mirror_spheres", 0, 5, 256, 256, "256x256"),
    ("horse_and_mug", 0, 16, 360, 180, "360x180"),  # the 16x16 path at a size the CPU renders in seconds
    ("marbles", 0, 16, 64, 64, "64x64"),
]

ROWS_8K = [0, 1, 777, 1500, 1919, 1920, 2000, 2345, 2800, 3333, 3838, 3839]


def xz(data):
    return lzma.compress(data, preset=9 | lzma.PRESET_EXTREME)


def ppm_md5(ref, img):
    with tempfile.NamedTemporaryFile(suffix=".ppm") as f:
        ref.L.ref_write_ppm(f.name.encode(), img.ctypes.data, img.shape[1], img.shape[0])
        return hashlib.md5(open(f.name, "rb").read()).hexdigest()


def main():
    os.makedirs(os.path.join(HERE, "scenes"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "images"), exist_ok=True)
    manifest = {"generator": "tests/golden/make_golden.py", "reference": "oracle/_ref/libref.so (unmodified sources)",
                "scenes": {}, "images": {}}
    for path in sorted(glob.glob(os.path.join(REF_INPUTS, "*.xml"))):
        scene_name = os.path.splitext(os.path.basename(path))[0]
        raw = open(path, "rb").read()
        with open(os.path.join(HERE, "scenes", scene_name + ".xml.xz"), "wb") as f:
            f.write(xz(raw))
        ref = H.RefScene(path)
        sc = ref.to_scene()
        orc = H.OracleScene(sc)
        assert orc.bvh_stats() == ref.bvh_stats(), scene_name
        manifest["scenes"][scene_name] = {
            "xml_sha256": hashlib.sha256(raw).hexdigest(),
            "loader_digest": sc.digest(),
            "counts": dict(zip(["vertices", "triangles", "meshes", "mesh_faces", "spheres", "materials", "lights", "cameras"], ref.counts)),
            "ref_bvh": dict(zip(["nodes", "leaves", "max_leaf", "max_depth"], ref.bvh_stats())),
            "cameras": [n for _, n in sc.cameras],
        }
        cases = [(ci, 1, 0, 0, None) for ci in range(len(sc.cameras))]
        cases += [(ci, aa, w, h, tag) for (s, ci, aa, w, h, tag) in AA_CASES if s == scene_name]
        for ci, aa, w, h, tag in cases:
            img, secs = ref.render(ci, aa, w, h)
            cam = sc.camera(ci, w or None, h or None)
            oimg, st = orc.render(cam, aa)
            rep = H.diff_report(img, oimg)
            assert rep["equal"] == rep["pixels"], (scene_name, ci, aa, rep)
            base = os.path.splitext(sc.cameras[ci][1])[0]
            key = f"{base}.aa{aa}" + (f".{tag}" if tag else "")
            with open(os.path.join(HERE, "images", key + ".rgb.xz"), "wb") as f:
                f.write(xz(img.tobytes()))
            manifest["images"][key] = {
                "scene": scene_name, "camera": ci, "aa": aa, "width": int(img.shape[1]), "height": int(img.shape[0]),
                "ppm_md5": ppm_md5(ref, img), "rgb_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
                "rays": {"primary": st.primary_rays, "reflection": st.reflection_rays, "shadow": st.shadow_rays,
                         "shadow_occluded": st.shadow_occluded},
                "ref_work": {"box_tests": st.box_tests, "tri_tests": st.tri_tests, "sphere_tests": st.sphere_tests},
                "ref_render_seconds_here": round(secs, 4),
            }
            print(key, img.shape, rep, f"{secs:.3f}s", flush=True)
        if scene_name == "horse_and_mug":
            # config 5: 7680x3840 output, 16x16 SSAA -> 122880x61440 sub-samples; selected output rows
            W, Hh, f = 7680, 3840, 16
            rows = np.zeros((len(ROWS_8K), W, 3), np.uint8)
            cam = sc.camera(0, W, Hh)
            for i, r in enumerate(ROWS_8K):
                secs, sub = ref.time_rows(0, f, W, Hh, r * f, 1, f, 0, keep=True)
                osub, _ = orc.render_rows(cam, f, r * f, 1, f)
                assert np.array_equal(sub, osub), r
                s = sub.astype(np.int64).reshape(f, W, f, 3).sum(axis=(0, 2))
                rows[i] = (s // (f * f)).astype(np.uint8)
                print("8k row", r, f"{secs:.2f}s", flush=True)
            key = "horse_and_mug_8k.aa16.rows"
            with open(os.path.join(HERE, "images", key + ".rgb.xz"), "wb") as fo:
                fo.write(xz(rows.tobytes()))
            manifest["images"][key] = {"scene": scene_name, "camera": 0, "aa": f, "width": W, "height": Hh,
                                       "rows": ROWS_8K, "rgb_sha256": hashlib.sha256(rows.tobytes()).hexdigest(),
                                       # known answers measured with the 64-bit-index build of the reference
                                       "full_frame_ppm_md5": "2be09d2f9bf003086286f36f4ee9cad0",
                                       "rays": {"primary": 7549747200, "reflection": 5797408973,
                                                "shadow": 12646772572}}
        orc.close()
        ref.close()
    full_key = "horse_and_mug_8k.aa16.full"
    if "--full-8k" in sys.argv:
        sc = H.load_scene_xml(H.golden_scene_path("horse_and_mug"))
        img, st = H.OracleScene(sc).render(sc.camera(0, 7680, 3840), 16)
        with tempfile.NamedTemporaryFile(suffix=".ppm") as f:
            H.write_ppm(f.name, img)
            md5 = hashlib.md5(open(f.name, "rb").read()).hexdigest()
        assert md5 == "2be09d2f9bf003086286f36f4ee9cad0", md5
        assert (st.primary_rays, st.reflection_rays, st.shadow_rays) == (7549747200, 5797408973, 12646772572)
        with open(os.path.join(HERE, "images", full_key + ".rgb.xz"), "wb") as fo:
            fo.write(lzma.compress(img.tobytes(), preset=6))
        manifest["images"][full_key] = {"scene": "horse_and_mug", "camera": 0, "aa": 16, "width": 7680, "height": 3840,
                                        "rgb_sha256": hashlib.sha256(img.tobytes()).hexdigest(), "ppm_md5": md5, "slow": True,
                                        "rays": {"primary": st.primary_rays, "reflection": st.reflection_rays,
                                                 "shadow": st.shadow_rays, "shadow_occluded": st.shadow_occluded}}
    else:  # keep the entry of an earlier --full-8k run
        try:
            old = json.load(open(os.path.join(HERE, "manifest.json")))["images"].get(full_key)
            if old:
                manifest["images"][full_key] = old
        except Exception:
            pass
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("done")


if __name__ == "__main__":
    main()
