"""The C-ABI library: loads without a GPU, exports every symbol include/rt_b200.h declares, its host-only
arithmetic (row-band partition) is right, and it fails loudly — not silently on a CPU path — when no GPU is present."""
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
    assert L.rt_abi_version() == 3


def test_band_partition_arithmetic():
    """rt_band_height / rt_part_rows / rt_part_bytes: a pure function of (camera, aa, world); the parts' bands tile the
    frame exactly once; the headline configuration deals single pixel rows like the reference (raytracer.cpp:353)."""
    L = H.rt_b200.cuda_lib()
    cam = H.RtCamera()
    for (w, h) in [(1, 1), (32, 32), (33, 31), (250, 190), (1440, 720), (7680, 3840)]:
        cam.image_width, cam.image_height = w, h
        for aa in (1, 2, 3, 4, 8, 16):
            for world in (1, 2, 3, 4, 8):
                bh = L.rt_band_height(C.byref(cam), aa, world)
                assert bh >= 1 and (aa % 8 != 0 or bh == 1)
                n_bands = (h + bh - 1) // bh
                rows = [L.rt_part_rows(C.byref(cam), aa, r, world) for r in range(world)]
                assert rows == [len(range(r, n_bands, world)) * bh for r in range(world)]
                assert sum(rows) == n_bands * bh
                assert all(L.rt_part_bytes(C.byref(cam), aa, r, world) == rows[r] * w * 3 for r in range(world))
                assert list(H.rt_b200.part_band_ids(h, bh, 0, world)) == list(range(0, n_bands, world))
    assert L.rt_part_rows(C.byref(cam), 1, 2, 2) == -1  # rank out of range
    cam.image_width, cam.image_height = 7680, 3840
    assert L.rt_band_height(C.byref(cam), 16, 8) == 1


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
