"""data/bench/cfgN.npz (read by bench.py's reference arm, numpy only) holds exactly what scenes.config(N) computes."""
import os

import numpy as np
import pytest
from _util import scenes


@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5, 6])
def test_fixture_matches_config(idx):
    sc, cam, fpd = scenes.config(idx)
    sc2, cam2, fpd2 = scenes.load_inputs(os.path.join(scenes.BENCH_DIR, f"cfg{idx}.npz"))
    assert fpd == fpd2 and len(sc.volumes) == len(sc2.volumes)
    for a, b in zip(sc.volumes, sc2.volumes):
        assert a.n == b.n and np.array_equal(a.grid, b.grid)
        for k in ("cube_min", "cube_max", "matrix", "inv_matrix"):
            assert np.array_equal(np.asarray(getattr(a, k), np.float32).view(np.uint32), np.asarray(getattr(b, k), np.float32).view(np.uint32)), k
    assert np.array_equal(np.asarray(sc.materials, np.float32).view(np.uint32), np.asarray(sc2.materials, np.float32).view(np.uint32))
    flat = lambda rows: [np.concatenate([np.ravel(np.asarray(x, np.float32)) for x in r]).view(np.uint32).tolist() for r in rows]
    for k in ("point_lights", "area_lights", "spot_lights"):
        assert flat(getattr(sc, k)) == flat(getattr(sc2, k)), k
    assert flat([sc.dir_light]) == flat([sc2.dir_light]) and sc.area_shadow_samples == sc2.area_shadow_samples
    assert (sc.sky is None) == (sc2.sky is None) and sc.sky_active == sc2.sky_active and sc.sky_contribution == sc2.sky_contribution
    if sc.sky is not None:
        assert np.array_equal(sc.sky, sc2.sky)
    for k in ("cam_pos", "top_left", "top_right", "bottom_left", "right", "up"):
        assert np.array_equal(np.asarray(getattr(cam, k), np.float32).view(np.uint32), np.asarray(getattr(cam2, k), np.float32).view(np.uint32)), k
    assert np.float32(cam.focal_distance) == np.float32(cam2.focal_distance) and np.float32(cam.defocus_jitter) == np.float32(cam2.defocus_jitter)
