#!/usr/bin/env python
"""Process wall time of `raytracer scene.xml` — this repository's CLI on the B200 against the reference's own
binary on the same box's host cores (README.md:1,8 quotes 0.452 s for horse_and_mug; raytracer.cpp:487-525).

    python tools/cli_wall.py [--runs 5] [--json out.json] [--md out.md] [scene[:aa] ...]

For every (scene, aa): both binaries are run `runs` times in scratch directories; reported are the wall time of
the whole process (fork to exit: CUDA context, XML, build, render, PPM files), the programs' own "Planted trees" /
"Rendered in" / "Total:" lines, and whether the PPM files are byte-identical.  aa = 2 is what the reference ships
(oracle/_ref/raytracer, unmodified); other factors use oracle/_ref/raytracer_aa (its main() with a run-time --aa).
The first run of our CLI in a fresh box is the cold start (driver + module load); it is reported separately.
"""
import argparse
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEFAULT = ["simple:1", "bunny:1", "horse_and_mug:1", "dragon_lowres:1", "mirror_spheres:1", "horse_and_mug:2", "car:1", "cornellbox:1"]


def run_once(cmd, cwd):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{cmd}: rc {p.returncode}: {p.stderr[-400:]}")
    out = {"wall_s": wall}
    for key, pat in (("planted_s", r"Planted trees in ([0-9.]+)"), ("rendered_s", r"Rendered in ([0-9.]+)"), ("total_s", r"Total: ([0-9.]+)")):
        m = re.search(pat, p.stdout)
        out[key] = float(m.group(1)) if m else None
    return out


def measure(scene, aa, runs, extra=()):
    import harness as H
    xml = H.golden_scene_path(scene)
    ours = os.path.join(H.PKG, "raytracer")
    ref = os.path.join(ROOT, "oracle", "_ref", "raytracer" if aa == 2 else "raytracer_aa")
    res = {"scene": scene, "aa": aa}
    work = tempfile.mkdtemp(prefix="cliwall_")
    try:
        da, db = os.path.join(work, "ours"), os.path.join(work, "ref")
        os.makedirs(da)
        os.makedirs(db)
        a = [run_once([ours, xml, "--aa", str(aa), *extra], da) for _ in range(runs)]
        res["ours"] = {k: statistics.median(x[k] for x in a) for k in a[0]}
        res["ours"]["first_run_wall_s"] = a[0]["wall_s"]
        res["ours"]["min_wall_s"] = min(x["wall_s"] for x in a)
        if os.path.exists(ref):
            cmd = [ref, xml] if aa == 2 else [ref, xml, "--aa", str(aa)]
            b = [run_once(cmd, db) for _ in range(max(1, min(runs, 3)))]
            res["reference"] = {k: statistics.median(x[k] for x in b) for k in b[0]}
            res["reference"]["min_wall_s"] = min(x["wall_s"] for x in b)
            names = sorted(os.listdir(db))
            res["files"] = names
            res["identical"] = bool(names) and sorted(os.listdir(da)) == names and all(
                open(os.path.join(da, n), "rb").read() == open(os.path.join(db, n), "rb").read() for n in names)
            res["speedup_wall"] = res["reference"]["wall_s"] / res["ours"]["wall_s"]
        else:
            res["reference"] = None
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return res


def table(rows, cores):
    out = [f"| scene | AA | cameras | ours: process wall s (median; first run) | ours: Planted / Rendered / Total | reference ({cores} host cores): process wall s | reference: Planted / Rendered / Total | wall speed-up | PPM files |",
           "|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        o, f = r["ours"], r.get("reference")
        fr = "-" if not f else f"{f['wall_s']:.3f}"
        ft = "-" if not f else f"{f['planted_s']:.3f} / {f['rendered_s']:.3f} / {f['total_s']:.3f}"
        sp = "-" if not f else f"{r['speedup_wall']:.2f}x"
        idt = "-" if not f else ("byte-identical" if r["identical"] else "DIFFERENT")
        out.append(f"| {r['scene']} | {r['aa']}x{r['aa']} | {len(r.get('files', [])) or '-'} | {o['wall_s']:.3f} ({o['first_run_wall_s']:.3f}) | "
                   f"{o['planted_s']:.3f} / {o['rendered_s']:.3f} / {o['total_s']:.3f} | {fr} | {ft} | {sp} | {idt} |")
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=DEFAULT)
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--json")
    ap.add_argument("--md")
    a = ap.parse_args()
    rows = []
    for c in a.cases:
        scene, _, aa = c.partition(":")
        rows.append(measure(scene, int(aa or 1), a.runs, extra=("--gpus", str(a.gpus)) if a.gpus > 1 else ()))
        print(json.dumps(rows[-1]), flush=True)
    md = table(rows, os.cpu_count())
    print(md)
    if a.json:
        json.dump({"host_cores": os.cpu_count(), "rows": rows}, open(a.json, "w"), indent=1)
    if a.md:
        open(a.md, "w").write(md + "\n")


if __name__ == "__main__":
    main()
