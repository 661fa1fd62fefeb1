#!/usr/bin/env python
"""Share of executed warp instructions (and of stall samples) per region of the render kernel, from the source page of
an `ncu --set full --import-source on` report built with -lineinfo:   python tools/ncu_regions.py report.ncu-rep"""
import csv
import os
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "raytracer-ceng477-graphics-hw-1_b200", "csrc")
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr and r and r[0].isdigit():
        try:
            out.append((cur, int(r[0]), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])))
        except Exception:
            pass
tot, ts = sum(o[2] for o in out), sum(o[3] for o in out)
dc = open(os.path.join(CSRC, "device_common.cuh")).read().splitlines()
rv = open(os.path.join(CSRC, "render_v2.cu")).read().splitlines()


def find(lines, pat):
    return next(i + 1 for i, l in enumerate(lines) if pat in l)


reg_dc = [("vector / division helpers (normalize, length, div3, dot)", 1, find(dc, "struct Ray")), ("make_ray", find(dc, "struct Ray"), find(dc, "// Cramer")),
          ("hit_triangle", find(dc, "// Cramer"), find(dc, "// raytracer.cpp:70-96")), ("hit_sphere", find(dc, "// raytracer.cpp:70-96"), find(dc, "// Box test.")),
          ("box test (slab)", find(dc, "// Box test."), find(dc, "// 256-bit read-only load")), ("node / primitive fetch", find(dc, "// 256-bit read-only load"), find(dc, "struct Counters")),
          ("closest_update / robust_visible", find(dc, "struct Counters"), find(dc, "// raytracer.cpp:101-126 bit for bit")),
          ("exact replay", find(dc, "// raytracer.cpp:101-126 bit for bit"), find(dc, "// parser.h:88-93")), ("quantise", find(dc, "// parser.h:88-93"), len(dc) + 1)]
reg_rv = [("start_primary", find(rv, "RT_DEV void start_primary"), find(rv, "template <bool FAR>")), ("traversal: setup", find(rv, "template <bool FAR>"), find(rv, "while (node != kSentinel)")),
          ("traversal: node loop", find(rv, "while (node != kSentinel)"), find(rv, "if (node < 0) {")), ("traversal: leaf loop", find(rv, "if (node < 0) {"), find(rv, "// reference visibility: a doubtful")),
          ("visibility check", find(rv, "// reference visibility: a doubtful"), find(rv, "// ---- consume the result")),
          ("consume / shade", find(rv, "// ---- consume the result"), find(rv, "RT_DEV unsigned char *pixel_ptr")), ("kernel body: items, accumulation, stores", find(rv, "RT_DEV unsigned char *pixel_ptr"), len(rv) + 1)]
c, cs = Counter(), Counter()
for f, l, n, s in out:
    regs = reg_dc if f == "device_common.cuh" else reg_rv if f == "render_v2.cu" else []
    name = next((nm for nm, a, b in regs if a <= l < b), f or "other")
    c[name] += n
    cs[name] += s
print(f"{rep}: {tot / 1e9:.2f} G warp instructions attributed to source lines (NOTE: line numbers are those of the build that was profiled)")
for k, v in c.most_common():
    if v * 1000 >= tot:
        print(f"{100 * v / tot:5.1f}% of instructions  {100 * cs[k] / max(ts, 1):5.1f}% of samples  {k}")
