"""The oracle against the round-2 digests made from the reference build (tools/make_golden_r2.py): analytic primitives
through the whole render path, a coloured directional light, area lights with 1 and 5 shadow samples, and equal-spp
accumulation at BASELINE.json's sample counts (16 spp for cfg 3 at full size, 64 spp for cfg 5 at 960x540)."""
import json
import os

import pytest
from _util import GOLDEN, R2_SMALL, frame_digest, params, scenes, small_config, vp

WANT = json.load(open(os.path.join(GOLDEN, "r2_digests.json")))


@pytest.mark.parametrize("name", sorted(R2_SMALL))
def test_small_scenes(oracle, name):
    sc, cam, fpd = R2_SMALL[name]()
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        acc, rgb8, st = s.render(cam, params(fpd))
    got = frame_digest(acc, rgb8, st)
    for k in got:
        assert got[k] == WANT[name][k], (name, k)
    assert st.shadow_rays > 0 and st.secondary_rays > 0


@pytest.mark.parametrize("name,idx,size,spp", [("cfg3_16spp", 3, None, 16), ("cfg5_64spp_960x540", 5, (960, 540), 64)])
def test_equal_spp_accumulation(oracle, name, idx, size, spp):
    sc, cam, fpd = scenes.config(idx) if size is None else small_config(idx, *size)
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        acc, rgb8, st = s.render(cam, params(fpd, spp=spp))
    got = frame_digest(acc, rgb8, st)
    for k in got:
        assert got[k] == WANT[name][k], (name, k)
