#!/usr/bin/env python3
"""Writes data/bench/cfg{1..6}.npz: the inputs of every bench / parity configuration as plain arrays (what the host mirror
computes: transforms, material table, lights, camera), so that bench.py's reference arm needs no library of this
repository.  tests/test_bench_inputs.py checks the files against scenes.config()."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")

if __name__ == "__main__":
    os.makedirs(scenes.BENCH_DIR, exist_ok=True)
    for idx in (1, 2, 3, 4, 5, 6):
        p = os.path.join(scenes.BENCH_DIR, f"cfg{idx}.npz")
        scenes.save_inputs(idx, p)
        print(p, os.path.getsize(p))
