"""include/nexus_b200_import.hpp (C++ host layer: Wavefront OBJ + MTL and Radiance .hdr readers) against the Python readers
(nexus_b200/obj.py, nexus_b200/hdr.py): same meshes, shading data, materials and pixels, value for value.  Host only, no GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

import test_hdr
import test_obj
from nexus_b200 import hdr, obj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "import_check")


def _build():
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), EXE + ".cpp",
                           "-L" + os.path.join(ROOT, "nexus_b200"), "-lnexus_b200", "-Wl,-rpath,$ORIGIN/../nexus_b200", "-o", EXE])


def test_cpp_readers_equal_the_python_readers(tmp_path):
    _build()
    cube = test_obj._write(tmp_path)
    sky = np.random.RandomState(3).uniform(0, 3, (9, 24, 3)).astype(np.float32)
    sky[2, 4:20] = (800.0, 600.0, 1.0)
    test_hdr._write(tmp_path / "sky.hdr", test_hdr._rgbe(sky), rle=True)
    r = subprocess.run([EXE, str(cube), str(tmp_path / "sky.hdr")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    want = obj.load_obj(cube)
    assert len(got["meshes"]) == len(want["meshes"]) and len(got["materials"]) == len(want["materials"])
    for g, w in zip(got["meshes"], want["meshes"]):
        assert g["name"] == w["name"] and g["material"] == w["material"]
        assert (np.array(g["triangles"], np.float32) == w["triangles"].ravel()).all()
        assert np.allclose(np.array(g["triangle_data"], np.float32), w["triangle_data"].ravel(), rtol=0, atol=1e-7)
    for g, w in zip(got["materials"], want["materials"]):
        assert np.allclose(g["baseColor"], w.baseColor) and np.allclose(g["emissionColor"], w.emissionColor) and g["intensity"] == w.intensity
        assert g["ior"] == pytest.approx(w.ior) and g["opacity"] == w.opacity and g["roughness"] == pytest.approx(w.roughness) and g["metalness"] == pytest.approx(w.metalness)
    px = hdr.load_hdr(tmp_path / "sky.hdr")
    assert (got["hdr"]["height"], got["hdr"]["width"]) == px.shape[:2]
    assert (np.array(got["hdr"]["rgba"], np.float32) == px.ravel()).all()
    # flat files and malformed files behave alike too
    test_hdr._write(tmp_path / "flat.hdr", test_hdr._rgbe(sky), rle=False)
    r2 = subprocess.run([EXE, str(cube), str(tmp_path / "flat.hdr")], capture_output=True, text=True)
    assert r2.returncode == 0 and (np.array(json.loads(r2.stdout)["hdr"]["rgba"], np.float32) == px.ravel()).all()
    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n")
    r3 = subprocess.run([EXE, str(tmp_path / "bad.obj"), str(tmp_path / "sky.hdr")], capture_output=True, text=True)
    assert r3.returncode == 1 and "out of range" in r3.stderr
    (tmp_path / "bad.hdr").write_bytes(b"P6\n1 1\n255\n...")
    r4 = subprocess.run([EXE, str(cube), str(tmp_path / "bad.hdr")], capture_output=True, text=True)
    assert r4.returncode == 1 and "not a Radiance" in r4.stderr
