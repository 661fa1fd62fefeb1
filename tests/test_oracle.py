"""The oracle (oracle/whitted_oracle.c) against the committed golden frames rendered by the unmodified reference.
CPU only.  This is what pins the checker: byte identity on every camera of every shipped scene, the supersampled
cases, selected rows of the full-size 8K 16x16 frame, and the reference's BVH statistics."""
import numpy as np
import pytest

import harness as H

M = H.manifest()
FRAMES = sorted(k for k, v in M["images"].items() if "rows" not in v and not v.get("slow"))


@pytest.mark.parametrize("key", FRAMES)
def test_oracle_matches_reference_frame(key):
    gold, m = H.golden_image(key)
    sc = H.golden_scene(m["scene"])
    orc = H.OracleScene(sc)
    img, st = orc.render(sc.camera(m["camera"], m["width"], m["height"]), m["aa"])
    orc.close()
    assert np.array_equal(img, gold), H.diff_report(gold, img)
    assert {"primary": st.primary_rays, "reflection": st.reflection_rays, "shadow": st.shadow_rays,
            "shadow_occluded": st.shadow_occluded} == m["rays"]


def test_oracle_matches_reference_8k_rows():
    """Config 5 (7680x3840, 16x16 SSAA): output rows built from the reference's own sub-samples."""
    gold, m = H.golden_image("horse_and_mug_8k.aa16.rows")
    sc = H.golden_scene("horse_and_mug")
    cam = sc.camera(0, m["width"], m["height"])
    orc = H.OracleScene(sc)
    f = m["aa"]
    for i in (0, 4, len(m["rows"]) - 1):  # three rows: top edge, mid-frame, bottom edge (~0.5 s each)
        r = m["rows"][i]
        sub, _ = orc.render_rows(cam, f, r * f, 1, f)
        row = (sub.astype(np.int64).reshape(f, m["width"], f, 3).sum(axis=(0, 2)) // (f * f)).astype(np.uint8)
        assert np.array_equal(row, gold[i]), r
    orc.close()


@pytest.mark.parametrize("scene", sorted(M["scenes"]))
def test_reference_bvh_statistics(scene):
    """Node / leaf / max-leaf / max-depth of the rebuilt reference tree (measured with the
    reference's own builder) — from the oracle AND from the product's rank builder (host-only C-ABI hook)."""
    import ctypes as C
    want = M["scenes"][scene]["ref_bvh"]
    sc = H.golden_scene(scene)
    orc = H.OracleScene(sc)
    assert dict(zip(["nodes", "leaves", "max_leaf", "max_depth"], orc.bvh_stats())) == want
    n = sc.desc.n_triangles + sc.desc.n_spheres
    ranks = np.zeros((8, n), np.uint32)
    stats = (C.c_int32 * 4)()
    L = H.rt_b200.cuda_lib()
    L.rt_host_reference_ranks.argtypes = [C.POINTER(H.RtSceneDesc), C.c_void_p, C.c_void_p]
    assert L.rt_host_reference_ranks(C.byref(sc.desc), ranks.ctypes.data, stats) == 0
    assert dict(zip(["nodes", "leaves", "max_leaf", "max_depth"], list(stats))) == want
    # the product's tie ranks are the oracle's traversal order, octant by octant
    assert np.array_equal(ranks, orc.visit_ranks())
    for o in range(8):
        assert np.array_equal(np.sort(ranks[o]), np.arange(n, dtype=np.uint32))
    orc.close()


@pytest.mark.skipif(not H.ref_available(), reason="oracle/_ref/libref.so not built (needs /root/reference)")
def test_oracle_matches_live_reference_on_unseen_configuration():
    """Beyond the fixtures: a resolution/AA combination nobody rendered before, reference vs oracle, live."""
    path = H.golden_scene_path("simple_reflectance")
    ref = H.RefScene(path)
    sc = ref.to_scene()
    want, _ = ref.render(0, 3, 101, 67)
    got, _ = H.OracleScene(sc).render(sc.camera(0, 101, 67), 3)
    ref.close()
    assert np.array_equal(want, got)


@pytest.mark.skipif(not H.ref_available(), reason="oracle/_ref/libref.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_oracle_matches_live_reference_on_seeded_scenes(seed, tmp_path):
    """Seeded triangle/sphere soups (flat boxes, shared edges, a degenerate triangle, mirrors, the camera inside a
    sphere for even seeds) written as XML, rendered by the unmodified reference and by the oracle: byte identity.
    Also checks the product's XML loader on a file the reference's loader parses."""
    sc = H.random_scene(seed, camera_inside_sphere=(seed % 2 == 0), max_depth=seed % 4)
    xml = str(tmp_path / f"random_{seed}.xml")
    H.scene_to_xml(sc, xml)
    ref = H.RefScene(xml)
    assert ref.to_scene().digest() == H.load_scene_xml(xml).digest()
    for aa in (1, 3):
        want, _ = ref.render(0, aa)
        got, _ = H.OracleScene(ref.to_scene()).render(ref.to_scene().camera(0), aa)
        assert np.array_equal(want, got), (seed, aa, H.diff_report(want, got))
    ref.close()
