"""The C ABI called from plain C (tests/abi_driver.c): what the Java shim of INTEGRATION.md does, without Python in between.

CPU: the driver compiles against include/node4j.h with gcc, links the in-tree library and — there being no device — gets
N4J_ERR_CUDA from node4j_create (no CPU fallback).  GPU: it runs the reference's torchdiffeq known-answer test
(MultiStepAdjointTest.java:174-232) and checks the golden vectors with the Java test's tolerances.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from neuralode4j_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_driver(tmp_path):
    libdir = os.path.dirname(_lib.lib_path())
    _lib.load()                                   # builds the library if it is stale
    exe = str(tmp_path / "abi_driver")
    cmd = ["gcc", "-O1", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_driver.c"),
           "-o", exe, "-L", libdir, "-lnode4j_b200", "-lm", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _golden_file(tmp_path):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
    vals = np.concatenate([np.asarray(g[k], np.float64).ravel() for k in ("ys_fwd", "dy0_fwd", "dtheta_fwd", "dt_fwd")])
    assert vals.size == 60
    p = tmp_path / "golden.txt"
    p.write_text("\n".join(repr(float(v)) for v in vals))
    return str(p)


def test_c_driver_compiles_and_fails_loudly_without_a_device(tmp_path):
    exe = _build_driver(tmp_path)
    if _lib.load().node4j_device_count() > 0:
        pytest.skip("a device is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "NO DEVICE -> N4J_ERR_CUDA" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_driver_reproduces_reference_golden_vectors(tmp_path):
    exe = _build_driver(tmp_path)
    r = subprocess.run([exe, _golden_file(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ABI DRIVER OK" in r.stdout, r.stdout + r.stderr
