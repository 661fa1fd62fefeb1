"""The C-ABI library: loads without a GPU, exports every symbol include/rt_b200.h declares, its host-only
arithmetic (tile partition) is right, and it fails loudly — not silently on a CPU path — when no GPU is present."""
import ctypes as C
import os
import re

import pytest

import harness as H


def declared_symbols():
    text = open(os.path.join(H.ROOT, "include", "rt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = H.rt_b200.cuda_lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert L.rt_abi_version() == 1


def test_tile_partition_arithmetic():
    L = H.rt_b200.cuda_lib()
    cam = H.RtCamera()
    for (w, h) in [(1, 1), (32, 32), (33, 31), (250, 190), (1440, 720), (7680, 3840)]:
        cam.image_width, cam.image_height = w, h
        for world in (1, 2, 3, 4, 8):
            tx, ty = H.rt_b200.tile_grid(w, h, world)  # includes the phantom column when tiles_x % world == 0
            assert tx in ((w + 31) // 32, (w + 31) // 32 + 1) and (tx % world != 0 or world == 1)
            tiles = tx * ty
            per = [L.rt_part_tiles(C.byref(cam), r, world) for r in range(world)]
            assert sum(per) == tiles
            assert per == [len(range(r, tiles, world)) for r in range(world)]
            assert all(L.rt_part_bytes(C.byref(cam), r, world) == per[r] * 32 * 32 * 3 for r in range(world))
    assert L.rt_part_tiles(C.byref(cam), 2, 2) == -1  # rank out of range


def test_no_cpu_fallback():
    """Without a GPU rt_scene_create must fail with RT_ERR_CUDA; with one this test has nothing to prove."""
    L = H.rt_b200.cuda_lib()
    if L.rt_device_count() > 0:
        pytest.skip("a GPU is present")
    sc = H.golden_scene("simple")
    with pytest.raises(H.rt_b200.RtError, match="error -2"):
        H.RayTracer(sc)
    # argument validation happens before any device work
    h = C.c_void_p()
    assert L.rt_scene_create(None, None, C.byref(h)) == -1
