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


def _glb_json(path, tmp_path):
    cube = test_obj._write(tmp_path)
    test_hdr._write(tmp_path / "tiny.hdr", test_hdr._rgbe(np.ones((1, 8, 3), np.float32)), rle=False)
    r = subprocess.run([EXE, str(cube), str(tmp_path / "tiny.hdr"), str(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)["glb"]


def _compare_glb(got, want):
    assert len(got["meshes"]) == len(want["meshes"]) and len(got["instances"]) == len(want["instances"]) and len(got["materials"]) == len(want["materials"])
    for g, w in zip(got["meshes"], want["meshes"]):
        assert g["name"] == w["name"] and g["material"] == w["material"]
        assert (np.array(g["triangles"], np.float32) == w["triangles"].ravel()).all()
        assert np.allclose(np.array(g["triangle_data"], np.float32), w["triangle_data"].ravel(), rtol=0, atol=1e-7)
    for g, w in zip(got["instances"], want["instances"]):
        assert g["mesh"] == w["mesh"] and np.allclose(np.array(g["matrix"], np.float32), w["matrix"].ravel(), rtol=1e-6, atol=1e-6)
    for g, w in zip(got["materials"], want["materials"]):
        ours = [*w.baseColor, w.opacity, w.metalness, w.roughness, *w.emissionColor, w.intensity, w.specularWeight, w.specularColor[0], w.ior, w.transmission, w.specularColor[2]]
        assert np.allclose(g, ours, rtol=1e-6, atol=1e-7)


def test_cpp_glb_reader_equals_the_python_reader(tmp_path):
    import test_gltf
    from nexus_b200 import gltf
    _build()
    test_gltf._two_quads_glb(tmp_path / "q.glb")
    got = _glb_json(tmp_path / "q.glb", tmp_path)
    _compare_glb(got, gltf.load_glb(tmp_path / "q.glb"))
    assert got["camera"] is False
    (tmp_path / "bad.glb").write_bytes(b"not a glb file at all....")
    cube = test_obj._write(tmp_path)
    r = subprocess.run([EXE, str(cube), str(tmp_path / "tiny.hdr"), str(tmp_path / "bad.glb")], capture_output=True, text=True)
    assert r.returncode == 1 and "not a binary glTF" in r.stderr


@pytest.mark.skipif(not os.path.exists("/root/reference/Nexus/assets/demo_scenes/cornell_box/cornell_box.glb"), reason="the reference's demo asset is only mounted in the build container")
def test_cpp_glb_reader_loads_the_reference_cornell_box(tmp_path):
    from nexus_b200 import gltf
    _build()
    ref = "/root/reference/Nexus/assets/demo_scenes/cornell_box/cornell_box.glb"
    got = _glb_json(ref, tmp_path)
    want = gltf.load_glb(ref)
    _compare_glb(got, want)
    assert len(got["meshes"]) == 8 and sum(len(m["triangles"]) // 9 for m in got["meshes"]) == 32


def test_cpp_glb_reader_decodes_embedded_textures(tmp_path):
    """The C++ reader decodes embedded PNG / JPEG texture images itself (include/nexus_b200_image.hpp; the reference: stb_image,
    IMGLoader.cpp:13-43) and assigns them like the Python reader: base colour and emissive sRGB, normal and metallic-roughness linear,
    one decode per (image, colour space)."""
    PIL = pytest.importorskip("PIL.Image")
    import io
    import test_gltf
    from nexus_b200 import gltf
    _build()
    rs = np.random.RandomState(5)
    px = rs.randint(0, 256, (4, 6, 4)).astype(np.uint8)
    buf = io.BytesIO(); PIL.fromarray(px, "RGBA").save(buf, format="PNG")
    png = buf.getvalue()
    smooth = np.clip(np.add.outer(np.arange(16) * 9.0, np.arange(24) * 6.0)[..., None] + np.array([0.0, 30.0, 60.0]), 0, 255).astype(np.uint8)
    buf = io.BytesIO(); PIL.fromarray(smooth, "RGB").save(buf, format="JPEG", quality=92, subsampling=0)
    jpg = buf.getvalue()
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    pad = b"\0" * ((4 - len(png) % 4) % 4)
    binary = pos.tobytes() + png + pad + jpg
    js = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
          "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "material": 0}]}],
          "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 1}}, "normalTexture": {"index": 0}, "emissiveTexture": {"index": 0}}],
          "textures": [{"source": 0}, {"source": 1}], "images": [{"bufferView": 1, "mimeType": "image/png"}, {"bufferView": 2, "mimeType": "image/jpeg"}],
          "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}],
          "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": pos.nbytes}, {"buffer": 0, "byteOffset": pos.nbytes, "byteLength": len(png)},
                          {"buffer": 0, "byteOffset": pos.nbytes + len(png) + len(pad), "byteLength": len(jpg)}],
          "buffers": [{"byteLength": len(binary)}]}
    test_gltf._write_glb(tmp_path / "t.glb", js, binary)
    got = _glb_json(tmp_path / "t.glb", tmp_path)
    want = gltf.load_glb(tmp_path / "t.glb")
    m = want["materials"][0]
    assert got["maps"][0] == [m.baseColorMap, m.emissiveMap, m.normalMap, m.metallicRoughnessMap]
    assert len(got["textures"]) == len(want["textures"]) == 3
    for g, (w_px, w_srgb) in zip(got["textures"], want["textures"]):
        assert (g["height"], g["width"]) == w_px.shape[:2] and g["srgb"] == w_srgb
        d = np.abs(np.array(g["rgba"], np.int32).reshape(w_px.shape) - w_px.astype(np.int32))
        assert d.max() <= (0 if g["width"] == 6 else 3)            # the PNG exactly, the JPEG within decoder rounding


DEMO_SCENES = "/root/reference/Nexus/assets/demo_scenes"


@pytest.mark.skipif(not os.path.isdir(DEMO_SCENES), reason="the reference's demo assets are only mounted in the build container")
def test_both_readers_load_every_demo_scene_of_the_reference(tmp_path):
    """All seven .glb demo scenes the reference ships (up to 3 M triangles, 21 embedded PNG / JPEG textures, progressive JPEGs among
    them): the Python reader (Pillow) and the C++ reader (its own decoders) produce the same counts and, for every texture, the
    same size, colour-space flag and pixels - PNG exactly, JPEG within decoder rounding."""
    import glob
    from nexus_b200 import gltf
    pytest.importorskip("PIL.Image")
    info = os.path.join(ROOT, "examples", "glb_info")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), info + ".cpp",
                           "-L" + os.path.join(ROOT, "nexus_b200"), "-lnexus_b200", "-Wl,-rpath,$ORIGIN/../nexus_b200", "-o", info])
    files = sorted(glob.glob(os.path.join(DEMO_SCENES, "*", "*.glb")))
    assert len(files) == 7
    textures = 0
    for path in files:
        want = gltf.load_glb(path)
        prefix = str(tmp_path / "tex")
        r = subprocess.run([info, path, prefix], capture_output=True, text=True)
        assert r.returncode == 0, (path, r.stderr)
        got = json.loads(r.stdout)
        assert got["meshes"] == len(want["meshes"]) and got["instances"] == len(want["instances"]) and got["materials"] == len(want["materials"]), path
        assert got["triangles"] == sum(len(m["triangles"]) for m in want["meshes"]), path
        assert len(got["textures"]) == len(want["textures"]), path
        js, _ = gltf._chunks(open(path, "rb").read())
        for k, (g, (w_px, w_srgb)) in enumerate(zip(got["textures"], want["textures"])):
            assert (g["height"], g["width"]) == w_px.shape[:2] and g["srgb"] == w_srgb, (path, k)
            px = np.fromfile(prefix + str(k) + ".rgba", np.uint8).reshape(w_px.shape).astype(np.int32)
            d = np.abs(px - w_px.astype(np.int32))
            # lossless images must be identical; JPEGs differ by decoder rounding (mostly 4:2:0 photographs: a few levels at chroma edges)
            if d.max() > 0:
                assert d.mean() <= 1.0 and np.percentile(d, 99.9) <= 12, (path, k, float(d.mean()), int(d.max()))
            textures += 1
    assert textures >= 40
