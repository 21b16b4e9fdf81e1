"""Minimal Wavefront OBJ (+ MTL) import into the scene description the host API consumes.

The reference reads every asset format through Assimp (src/Assets/OBJLoader.cpp:420-446: Triangulate | FlipUVs | CalcTangentSpace;
one Mesh per material group, one identity MeshInstance per mesh, materials from the keys of OBJLoader.cpp:96-119).  Assimp is not
available here; this reader covers what an OBJ can carry:

  geometry   v / vt / vn, faces `v`, `v/vt`, `v//vn`, `v/vt/vn` with positive or negative (relative) indices, polygons
             triangulated as fans; faces are grouped by `usemtl` into one mesh per material (as Assimp splits them);
             texture coordinates get the reference's V flip (aiProcess_FlipUVs); missing normals -> geometric normals
  materials  MTL `Kd` -> baseColor, `Ke` -> emissionColor (intensity 1, OBJLoader.cpp:113-118), `Ni` -> ior, `d` / `Tr` -> opacity,
             and the PBR extension Assimp reads: `Pr` -> roughness, `Pm` -> metalness.  Texture statements (`map_*`) are rejected
             unless `ignore_maps=True` (no image decoder here, as for .glb)

Pure host code: numpy only, no GPU.
"""
import os

import numpy as np

from . import Camera, Material, RenderSettings


class ObjError(ValueError):
    pass


def _load_mtl(path, ignore_maps):
    mats, cur = {}, None
    with open(path, "r", errors="replace") as f:
        for ln, line in enumerate(f, 1):
            tok = line.split("#", 1)[0].split()
            if not tok:
                continue
            key, val = tok[0], tok[1:]
            if key == "newmtl":
                cur = Material()
                mats[" ".join(val)] = cur
                continue
            if cur is None:
                raise ObjError(f"{path}:{ln}: statement before the first newmtl")
            try:
                if key == "Kd":
                    cur.baseColor = tuple(float(x) for x in val[:3])
                elif key == "Ke":
                    e = tuple(float(x) for x in val[:3])
                    if any(e):
                        cur.emissionColor, cur.intensity = e, 1.0
                elif key == "Ni":
                    cur.ior = float(val[0])
                elif key == "d":
                    cur.opacity = float(val[0])
                elif key == "Tr":
                    cur.opacity = 1.0 - float(val[0])
                elif key == "Pr":
                    cur.roughness = float(val[0])
                elif key == "Pm":
                    cur.metalness = float(val[0])
                elif key.startswith("map_") or key in ("bump", "disp", "norm"):
                    if not ignore_maps:
                        raise ObjError(f"{path}:{ln}: texture maps need an image decoder (pass ignore_maps=True to drop them)")
            except ObjError:
                raise
            except (ValueError, IndexError):
                raise ObjError(f"{path}:{ln}: malformed '{key}' statement") from None
    return mats


def load_obj(path, path_length=10, ignore_maps=False):
    """Reads an .obj file (and the .mtl files it names) into the scene description `scenes.build` instantiates."""
    v, vt, vn = [], [], []
    groups, order = {}, []              # material name -> list of (3 x (vi, ti, ni)) triangles
    mtl, current = {}, None
    base = os.path.dirname(os.path.abspath(path))

    def index(tok, count, what, ln):
        if tok == "":
            return -1
        i = int(tok)
        j = i - 1 if i > 0 else count + i
        if i == 0 or not (0 <= j < count):
            raise ObjError(f"{path}:{ln}: {what} index {i} out of range")
        return j

    with open(path, "r", errors="replace") as f:
        for ln, line in enumerate(f, 1):
            tok = line.split("#", 1)[0].split()
            if not tok:
                continue
            key, val = tok[0], tok[1:]
            try:
                if key == "v":
                    v.append([float(x) for x in val[:3]])
                elif key == "vt":
                    vt.append([float(val[0]), float(val[1]) if len(val) > 1 else 0.0])
                elif key == "vn":
                    vn.append([float(x) for x in val[:3]])
                elif key == "f":
                    if len(val) < 3:
                        raise ObjError(f"{path}:{ln}: a face needs at least three vertices")
                    corners = []
                    for c in val:
                        p = (c.split("/") + ["", ""])[:3]
                        corners.append((index(p[0], len(v), "vertex", ln), index(p[1], len(vt), "texture", ln), index(p[2], len(vn), "normal", ln)))
                    if current not in groups:
                        groups[current] = []
                        order.append(current)
                    for k in range(1, len(corners) - 1):                       # fan, like aiProcess_Triangulate on convex polygons
                        groups[current].append((corners[0], corners[k], corners[k + 1]))
                elif key == "usemtl":
                    current = " ".join(val)
                elif key == "mtllib":
                    for name in val:
                        mtl.update(_load_mtl(os.path.join(base, name), ignore_maps))
            except ObjError:
                raise
            except (ValueError, IndexError):
                raise ObjError(f"{path}:{ln}: malformed '{key}' statement") from None
    if not order:
        raise ObjError("the file contains no faces")
    V = np.asarray(v, np.float32).reshape(-1, 3)
    VT = np.asarray(vt, np.float32).reshape(-1, 2)
    VN = np.asarray(vn, np.float32).reshape(-1, 3)
    materials, material_of = [], {}
    meshes, instances = [], []
    for name in order:
        if name not in material_of:
            if name is not None and name not in mtl:
                raise ObjError(f"material '{name}' is not defined by any mtllib")
            material_of[name] = len(materials)
            materials.append(mtl[name] if name is not None else Material())
        idx = np.asarray(groups[name], np.int64).reshape(-1, 3, 3)               # (triangle, corner, (v, vt, vn))
        tris = V[idx[:, :, 0]].reshape(-1, 9)
        n = len(tris)
        data = np.zeros((n, 24), np.float32)
        g = np.cross(tris[:, 3:6] - tris[:, 0:3], tris[:, 6:9] - tris[:, 0:3])
        g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-30)
        has_n = (idx[:, :, 2] >= 0).all(axis=1)
        normals = np.repeat(g[:, None, :], 3, axis=1)
        if len(VN) and has_n.any():
            normals[has_n] = VN[idx[has_n][:, :, 2]]
        data[:, 0:9] = normals.reshape(n, 9)
        has_t = (idx[:, :, 1] >= 0).all(axis=1)
        if len(VT) and has_t.any():
            uv = VT[idx[has_t][:, :, 1]].copy()
            uv[:, :, 1] = 1.0 - uv[:, :, 1]                                      # aiProcess_FlipUVs
            data[has_t, 18:24] = uv.reshape(-1, 6)
        meshes.append({"name": f"{os.path.basename(path)}.{name or 'default'}", "triangles": np.ascontiguousarray(tris, np.float32),
                       "triangle_data": data, "material": material_of[name]})
        instances.append({"mesh": len(meshes) - 1, "material": -1, "matrix": np.eye(4, dtype=np.float32)})
    return {"name": str(path), "meshes": meshes, "instances": instances, "materials": materials, "textures": [], "lights": [],
            "camera": Camera(), "settings": RenderSettings(useMIS=True, pathLength=path_length)}
