#!/usr/bin/env python
"""How much slower is a GPU on its 1/N share of the frame than on 1/N of the whole frame's time?  (ONE GPU renders the
N parts one after the other.)  Sweeps the band height and the strip width of the 16x16 kernel via rt_set_partition.
    python tools/partition_experiment.py [world] [width height aa]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import torch  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w, h, aa = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (7680, 3840, 16)
sc = H.golden_scene("horse_and_mug")
cam = sc.camera(0, w, h)
rt = H.RayTracer(sc)
L = H.rt_b200.cuda_lib()
buf = torch.zeros(w * h * 3 + (1 << 20), dtype=torch.uint8, device="cuda")


def part_ms(rank, n):
    best = 1e30
    for _ in range(2):
        st = rt.render_part(cam, aa, rank, n, buf.data_ptr(), 0, want_stats=True)
        best = min(best, st.ms_render)
    return best


if len(sys.argv) > 5 and sys.argv[5] == "blocks":  # block shape sweep (rt_set_block_width) at the default band height / run length
    L.rt_set_block_width(32)
    full = part_ms(0, 1)
    for n in (1, 2, 4, world):
        for bw in (32, 64, 128, 256):
            L.rt_set_block_width(bw)
            ms = [part_ms(r, n) for r in range(n)]
            print(f"world {n} block {bw:3d}x{1024 // bw:<3d}: max {max(ms):.2f} mean {sum(ms) / n:.2f} ms; ideal {full / n:.2f}; "
                  f"efficiency if the slowest part decides {full / n / max(ms):.4f}, mean part {full / n / (sum(ms) / n):.4f}", flush=True)
    rt.close()
    sys.exit(0)
for strip in (0, 8, 4, 2):
    L.rt_set_partition(0, strip)
    full = part_ms(0, 1)
    print(f"full frame, strip width {strip or 'default'}: {full:.2f} ms", flush=True)
L.rt_set_partition(0, 0)
full = part_ms(0, 1)
for band_rows in (1, 4):
    for strip in (8, 4, 2):
        L.rt_set_partition(band_rows, strip)
        ms = [part_ms(r, world) for r in range(world)]
        print(f"world {world} band_rows {band_rows:2d} strip {strip or 'default':>7}: max {max(ms):.2f} mean {sum(ms) / world:.2f} ms; ideal {full / world:.2f}; "
              f"efficiency if the slowest part decides {full / world / max(ms):.4f}, mean part {full / world / (sum(ms) / world):.4f}", flush=True)
rt.close()
