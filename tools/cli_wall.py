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


def probe(runs=5):
    """Wall time of a process that only brings CUDA up (`raytracer --probe`: driver, context, kernel load)."""
    import harness as H
    ours = os.path.join(H.PKG, "raytracer")
    t = []
    for _ in range(runs):
        t0 = time.perf_counter()
        subprocess.run([ours, "--probe"], capture_output=True)
        t.append(time.perf_counter() - t0)
    return {"median_s": statistics.median(t), "min_s": min(t), "max_s": max(t)}


def batch(aa=1, runs=2):
    """All shipped scenes: ONE process of this CLI (`raytracer a.xml b.xml ...`: one CUDA start-up) against one process of
    the reference's binary per scene."""
    import harness as H
    scenes = sorted(H.manifest()["scenes"])
    xmls = [H.golden_scene_path(s) for s in scenes]
    ours = os.path.join(H.PKG, "raytracer")
    ref = os.path.join(ROOT, "oracle", "_ref", "raytracer_aa")
    work = tempfile.mkdtemp(prefix="cliwall_batch_")
    try:
        da, db = os.path.join(work, "ours"), os.path.join(work, "ref")
        os.makedirs(da)
        os.makedirs(db)
        t = []
        for _ in range(runs):
            t0 = time.perf_counter()
            p = subprocess.run([ours, *xmls, "--aa", str(aa)], cwd=da, capture_output=True, text=True)
            t.append(time.perf_counter() - t0)
            if p.returncode != 0:
                raise RuntimeError(p.stderr[-400:])
        own_total = sum(float(x) for x in re.findall(r"Total: ([0-9.]+)", p.stdout))
        out = {"scenes": len(scenes), "aa": aa, "ours_one_process_wall_s": min(t), "ours_sum_of_totals_s": own_total}
        if os.path.exists(ref):
            t0 = time.perf_counter()
            for x in xmls:
                subprocess.run([ref, x, "--aa", str(aa)], cwd=db, capture_output=True, check=True)
            out["reference_processes_wall_s"] = time.perf_counter() - t0
            names = sorted(os.listdir(db))
            out["files"] = len(names)
            out["identical"] = sorted(os.listdir(da)) == names and all(open(os.path.join(da, n), "rb").read() == open(os.path.join(db, n), "rb").read() for n in names)
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


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
    pr = probe()
    md = table(rows, os.cpu_count())
    md += (f"\n\nCUDA start-up alone on this box (`raytracer --probe`: driver + context + kernel load, no scene): median {pr['median_s']:.3f} s "
           f"(min {pr['min_s']:.3f}, max {pr['max_s']:.3f}) — included in every \"ours: process wall\" above.")
    bt = batch() if a.gpus == 1 else None
    if bt:
        md += (f"\n\nAll {bt['scenes']} shipped scenes, no AA ({bt.get('files', '?')} PPM files): ONE process of this CLI {bt['ours_one_process_wall_s']:.3f} s wall "
               f"(its own \"Total:\" lines add up to {bt['ours_sum_of_totals_s']:.3f} s) against {bt.get('reference_processes_wall_s', float('nan')):.3f} s for the "
               f"reference's binary run once per scene; files {'byte-identical' if bt.get('identical') else 'DIFFERENT'}.")
    print(md)
    if a.json:
        json.dump({"host_cores": os.cpu_count(), "cuda_startup_probe": pr, "batch": bt, "rows": rows}, open(a.json, "w"), indent=1)
    if a.md:
        open(a.md, "w").write(md + "\n")


if __name__ == "__main__":
    main()
