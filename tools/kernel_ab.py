#!/usr/bin/env python
"""A/B timing of builds of the CUDA library (tools/build_variants.sh) and of run-time options on fixed cases.

    python tools/kernel_ab.py [--variants base,nodiv3,...] [--json out.json] [scene:width:height:aa ...]

Every variant runs in its own process (RT_B200_LIB selects the library); reported is the best render-kernel time of
6 frames (CUDA events), the rays/s, and whether the frame equals the first variant's byte for byte."""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raytracer-ceng477-graphics-hw-1_b200")
DEFAULT_CASES = ["horse_and_mug:3840:1920:16", "horse_and_mug:1440:720:1", "marbles:2048:2048:4", "car:2048:1536:8", "dragon_lowres:1600:1600:2"]


def child(cases, opts):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    for case in cases:
        name, w, h, aa = case.split(":")
        sc = H.golden_scene(name)
        cam = sc.camera(0, int(w), int(h))
        rt = H.RayTracer(sc, **opts)
        best = 1e30
        for _ in range(6):
            img = rt.render(cam, int(aa))
            best = min(best, rt.last_stats.ms_render)
        st = rt.last_stats
        print(json.dumps({"case": case, "ms": best, "mrays_s": st.total_rays / best / 1e3, "rays": st.total_rays,
                          "sha": hashlib.sha256(img.tobytes()).hexdigest()[:16], "builder": rt.info().builder}), flush=True)
        rt.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=DEFAULT_CASES)
    ap.add_argument("--variants", default="base")
    ap.add_argument("--json")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--opts", default="{}")
    a = ap.parse_args()
    if a.child:
        return child(a.cases, json.loads(a.opts))
    results = {}
    for v in a.variants.split(","):
        name, _, opts = v.partition("@")  # name@{"refill_threshold":8}
        env = dict(os.environ)
        lib = os.path.join(PKG, "libwhitted_b200.so") if name == "product" else os.path.join(PKG, "ab", name + ".so")
        env["RT_B200_LIB"] = lib
        p = subprocess.run([sys.executable, __file__, "--child", "--opts", opts or "{}", *a.cases], env=env, capture_output=True, text=True)
        if p.returncode != 0:
            print(v, "FAILED", p.stderr[-500:])
            continue
        results[v] = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    first = next(iter(results.values()), [])
    print(f"{'case':34s} " + " ".join(f"{v[:22]:>22s}" for v in results))
    for i, c in enumerate(first):
        row = []
        for v, rs in results.items():
            r = rs[i]
            row.append(f"{r['ms']:9.3f} ms {r['mrays_s'] / 1e3:6.2f}G{'' if r['sha'] == c['sha'] else ' !!'}")
        print(f"{c['case']:34s} " + " ".join(f"{x:>22s}" for x in row))
    if a.json:
        json.dump(results, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
