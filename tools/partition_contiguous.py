#!/usr/bin/env python
"""Where does a 1/8 part of the frame lose time against 1/8 of the whole frame's time?  ONE GPU renders the 8 parts one after
the other with different band heights (rt_set_partition) and pixel-block shapes (rt_set_block_width): contiguous eighths cost
the same as the whole frame (no per-launch overhead), single interleaved rows cost 1.6 % more (the rows a part owns are 8
image rows apart: the warps that run together cover 8x the image area, the L1 working set grows)."""
import os, sys
sys.path.insert(0, "tests")
import harness as H, torch
w, h, aa, world = 7680, 3840, 16, 8
sc = H.golden_scene("horse_and_mug"); cam = sc.camera(0, w, h); rt = H.RayTracer(sc); L = H.rt_b200.cuda_lib()
buf = torch.zeros(w * h * 3 + (1 << 20), dtype=torch.uint8, device="cuda")
def part_ms(rank, n):
    best = 1e30
    for _ in range(2):
        st = rt.render_part(cam, aa, rank, n, buf.data_ptr(), 0, want_stats=True); best = min(best, st.ms_render)
    return best
for rows, bw in ((0, 32), (480, 32), (60, 32), (4, 256), (8, 128), (16, 64), (32, 32), (8, 32), (2, 512)):
    L.rt_set_partition(rows, 0)
    L.rt_set_block_width(bw)
    full = part_ms(0, 1)
    ms = [part_ms(r, world) for r in range(world)]
    print(f"band_rows {rows or 1} block width {bw}: max part {max(ms):.2f} (ideal {full / world:.2f}, efficiency {full / world / max(ms):.4f}); full {full:.2f}; parts {[round(x, 2) for x in ms]} sum {sum(ms):.2f} (sum - full = {sum(ms) - full:.2f} ms)", flush=True)
