#!/usr/bin/env python3
"""Round-2 golden digests, from the REFERENCE ITSELF (oracle/_ref/libvpref.so), build container only:
  * the small parity scenes of tests/_util.py (analytic primitives through the whole render path, a coloured directional
    light, area lights with 1 and 5 shadow samples): SHA-256 of accumulator and RGB8 + ray counters;
  * equal-spp accumulation at BASELINE.json's sample counts: cfg 3 at 1920x1080 with 16 spp, cfg 5 at 960x540 with 64 spp.
Writes tests/golden/r2_digests.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import GOLDEN, R2_SMALL, frame_digest, params, ref_lib, scenes, small_config, vp  # noqa: E402


def main():
    ref = ref_lib()
    assert ref is not None, "build oracle/_ref first (python oracle/ref/build_ref.py)"
    out = {}
    for name, make in R2_SMALL.items():
        sc, cam, fpd = make()
        with vp.Session(ref) as s:
            s.set_scene(sc)
            acc, rgb8, st = s.render(cam, params(fpd))
        out[name] = dict(width=fpd["width"], height=fpd["height"], spp=fpd["spp"], max_bounces=fpd["max_bounces"], **frame_digest(acc, rgb8, st))
        print(name, out[name]["stats"], flush=True)
    for name, idx, size, spp in (("cfg3_16spp", 3, None, 16), ("cfg5_64spp_960x540", 5, (960, 540), 64)):
        sc, cam, fpd = scenes.config(idx) if size is None else small_config(idx, *size)
        with vp.Session(ref) as s:
            s.set_scene(sc)
            acc, rgb8, st = s.render(cam, params(fpd, spp=spp))
        out[name] = dict(width=fpd["width"], height=fpd["height"], spp=spp, max_bounces=fpd["max_bounces"], **frame_digest(acc, rgb8, st))
        print(name, out[name]["stats"], flush=True)
    with open(os.path.join(GOLDEN, "r2_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
