"""Asset import: the minimal Wavefront OBJ + MTL reader (nexus_b200/obj.py) that stands in for the reference's
Assimp import of .obj files (src/Assets/OBJLoader.cpp:96-119, 420-446).  Host logic: only the last test needs a GPU."""
import numpy as np
import pytest

import nexus_b200 as nx
from nexus_b200 import obj, scenes

OBJ = """# a unit cube: quads, mixed index forms, relative indices, two materials, one polygon without normals
mtllib cube.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0 0 1
v 1 0 1
v 1 1 1
v 0 1 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 -1
vn 0 0 1
usemtl red
f 1/1/1 4/4/1 3/3/1 2/2/1
f 5/1/2 6/2/2 7/3/2 8/4/2
usemtl lamp
f 1//1 2//1 6//1
f -8 -4 -3 -7
usemtl red
f 2 3 7 6
"""
MTL = """newmtl red
Kd 0.8 0.1 0.05
Ni 1.45
d 0.75
Pr 0.6
Pm 0.2
newmtl lamp
Kd 0 0 0
Ke 4 3 2
"""


def _write(tmp_path, obj_text=OBJ, mtl_text=MTL):
    (tmp_path / "cube.obj").write_text(obj_text)
    (tmp_path / "cube.mtl").write_text(mtl_text)
    return tmp_path / "cube.obj"


def test_obj_reader_groups_by_material_and_triangulates(tmp_path):
    d = obj.load_obj(_write(tmp_path))
    assert [m["name"].split(".")[-1] for m in d["meshes"]] == ["red", "lamp"] and len(d["instances"]) == 2
    red, lamp = d["meshes"]
    assert red["triangles"].shape == (6, 9) and lamp["triangles"].shape == (3, 9)          # 3 quads -> 6, triangle + quad -> 3
    assert (d["instances"][0]["matrix"] == np.eye(4)).all() and red["material"] == 0 and lamp["material"] == 1
    # first quad 1 4 3 2 -> fan (1, 4, 3), (1, 3, 2)
    assert (red["triangles"][0] == [0, 0, 0, 0, 1, 0, 1, 1, 0]).all() and (red["triangles"][1] == [0, 0, 0, 1, 1, 0, 1, 0, 0]).all()
    td = red["triangle_data"]
    assert (td[0, 0:9].reshape(3, 3) == [0, 0, -1]).all() and (td[2, 0:9].reshape(3, 3) == [0, 0, 1]).all()
    assert (td[0, 18:24] == [0, 1, 0, 0, 1, 0]).all()                                        # uv (0,0) (0,1) (1,1) with V flipped
    # the last red quad (2 3 7 6) has no normals: geometric normal +x at all three corners, no texture coordinates
    assert np.allclose(td[4, 0:9].reshape(3, 3), [1, 0, 0]) and (td[4, 18:24] == 0).all()
    # relative indices: -8 -4 -3 -7 = 1 5 6 2 -> the y = 0 face
    assert (lamp["triangles"][1:, 1::3] == 0).all()
    a, b = d["materials"]
    assert np.allclose(a.baseColor, (0.8, 0.1, 0.05)) and a.ior == pytest.approx(1.45) and a.opacity == 0.75 and a.roughness == 0.6 and a.metalness == 0.2
    assert a.intensity == 0.0 and np.allclose(b.emissionColor, (4, 3, 2)) and b.intensity == 1.0


def test_obj_reader_rejects_what_it_cannot_represent(tmp_path):
    for text, what in (("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n", "out of range"), ("v 0 0 0\n", "no faces"), ("v 0 0 0\nv 1 0 0\nf 1 2\n", "three"),
                       ("usemtl ghost\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", "not defined"), ("v 0 0 zero\n", "malformed")):
        with pytest.raises(obj.ObjError, match=what):
            obj.load_obj(_write(tmp_path, text))
    with pytest.raises(obj.ObjError, match="decoder"):
        obj.load_obj(_write(tmp_path, mtl_text=MTL + "map_Kd wood.png\n"))
    assert len(obj.load_obj(_write(tmp_path, mtl_text=MTL + "map_Kd wood.png\n"), ignore_maps=True)["materials"]) == 2
    d = obj.load_obj(_write(tmp_path, "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n"))                  # no material at all: the default one
    assert len(d["materials"]) == 1 and d["meshes"][0]["material"] == 0


@pytest.mark.gpu
def test_obj_scene_renders_like_the_same_scene_built_directly(tmp_path):
    d = obj.load_obj(_write(tmp_path), path_length=3)
    # seen from below and in front: the emissive y = 0 face is in view, and a grey sky lights the rest
    d["camera"] = nx.Camera(position=(0.5, -2.0, 4.0), forward=(0.0, 0.5547002, -0.8320503), horizontalFOV=40.0)
    d["settings"].backgroundColor = (0.4, 0.5, 0.6)
    d["materials"][0].opacity = 1.0
    res = (96, 64)
    ctx = nx.Context(0)
    a = scenes.build(ctx, d, res)
    direct = {k: v for k, v in d.items()}
    direct["instances"] = [{"mesh": i["mesh"], "material": -1, "position": (0, 0, 0), "rotation": (0, 0, 0), "scale": (1, 1, 1)} for i in d["instances"]]
    b = scenes.build(ctx, direct, res)
    pa, pb = nx.PathTracer(ctx, res), nx.PathTracer(ctx, res)
    pa.Render(a, frames=8, firstFrame=1); pb.Render(b, frames=8, firstFrame=1)
    ia, ib = pa.ReadAccumulation(), pb.ReadAccumulation()
    assert ia.mean() > 0 and np.allclose(ia, ib, rtol=1e-4, atol=1e-5)
    o, dd = scenes.camera_rays(d["camera"], res)
    hits = a.TraceClosest(nx.make_rays(o, dd))
    assert (hits["prim"] != 0xffffffff).mean() > 0.03 and set(np.unique(hits["instance"][hits["prim"] != 0xffffffff])) <= {0, 1}
    pa.close(); pb.close(); a.close(); b.close(); ctx.close()
