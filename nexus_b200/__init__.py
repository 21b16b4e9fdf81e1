"""nexus_b200 — B200-native (sm_100a) implementation of the StokastX/Nexus wavefront path-tracing hot path.

Python mirror of the reference's host API (same class and method names, reference file:line in each docstring), a thin
layer over the C ABI in include/nexus_b200.h.  All computation happens in libnexus_b200.so on the GPU; nothing here has
a CPU fallback.
"""
import ctypes as C

import numpy as np

from ._capi import (Aabb, BuildConfig, BuildMetrics, Bvh2, Bvh8, CameraPod, FrameStats, KernelProfile, LightPod, MaterialPod, NexusError,
                    RenderSettingsPod, check, lib)

__all__ = ["Context", "Scene", "AssetManager", "Material", "Light", "Camera", "RenderSettings", "PathTracer", "MeshInstance",
           "BuildBVH2", "BuildBVH8", "BenchmarkBuild", "BVH2", "BVH8", "NexusError", "write_pfm", "write_exr"]

RAY_DTYPE = np.dtype([("origin", np.float32, 3), ("tmax", np.float32), ("direction", np.float32, 3), ("pad", np.uint32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.uint32), ("instance", np.uint32)])
MISS_T = np.float32(1.0e30)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU.  Replaces the reference's process-global device state (src/Cuda/PathTracer/PathTracer.cu:21-37)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().nx_ctx_create(int(device), C.byref(self._h))
        if rc < 0:
            raise NexusError(f"nx_ctx_create(device={device}) failed with status {rc}: no usable CUDA device (no CPU fallback exists)")
        self.device = int(device)

    def close(self):
        if self._h:
            lib().nx_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self._h, lib().nx_ctx_synchronize(self._h), "synchronize")

    def SetTraceTuning(self, tri_lanes, inst_lanes):
        check(self._h, lib().nx_ctx_set_trace_tuning(self._h, C.c_uint32(tri_lanes), C.c_uint32(inst_lanes)), "SetTraceTuning")

    def SetSceneCollapse(self, collapse, maxLeafPrims=0):
        """BVH2 -> BVH8 collapse used for the BLASes / TLAS of scenes created afterwards (default: SAH-optimal, 2 primitives per
        leaf; COLLAPSE_REFERENCE_GPU gives trees identical to the reference's NexusBVH)."""
        check(self._h, lib().nx_ctx_set_scene_collapse(self._h, C.c_int(int(collapse)), C.c_int(int(maxLeafPrims))), "SetSceneCollapse")

    def SetInstanceMerging(self, enabled):
        """On (default): instances whose mesh no other instance uses, and that were never moved, share one world-space BLAS.
        Applies to scenes updated afterwards.  Hits are unchanged."""
        check(self._h, lib().nx_ctx_set_instance_merging(self._h, C.c_int(int(enabled))), "SetInstanceMerging")

    def SetTlasRefit(self, enabled):
        """Scene.Update after instances moved: False (default, the reference's behaviour) rebuilds the TLAS, True refits it whenever the
        set of TLAS entries is unchanged.  Same hits either way."""
        check(self._h, lib().nx_ctx_set_tlas_refit(self._h, C.c_int(int(enabled))), "SetTlasRefit")

    def SetTraceMode(self, mode):
        """'lane' (default: one ray per lane, loop specialised for the scene kind), 'general' (the same loop unspecialised), 'pool' (64
        rays per warp in shared memory, lanes take rays by kind of work) or 'duo' (two rays per lane).  Same hits in all of them."""
        m = {"lane": 0, "pool": 1, "duo": 2, "general": 3}.get(mode, mode)
        check(self._h, lib().nx_ctx_set_trace_mode(self._h, C.c_int(int(m))), "SetTraceMode")

    def SetPoolTuning(self, node, tri, inst, fetch, any_hit=False):
        check(self._h, lib().nx_ctx_set_pool_tuning(self._h, C.c_int(int(any_hit)), C.c_uint32(node), C.c_uint32(tri), C.c_uint32(inst), C.c_uint32(fetch)), "SetPoolTuning")

    def SetStackLimit(self, entries):
        check(self._h, lib().nx_ctx_set_stack_limit(self._h, C.c_uint32(entries)), "SetStackLimit")

    def SetSphereCull(self, enabled):
        check(self._h, lib().nx_ctx_set_sphere_cull(self._h, C.c_int(int(enabled))), "SetSphereCull")

    @property
    def sm_count(self):
        return lib().nx_ctx_sm_count(self._h)

    # device memory (N/Device/CudaMemory.h)
    def malloc(self, nbytes):
        p = C.c_void_p()
        check(self._h, lib().nx_malloc(self._h, C.c_size_t(nbytes), C.byref(p)), "nx_malloc")
        return p.value

    def free(self, dev):
        check(self._h, lib().nx_free(self._h, C.c_void_p(dev)), "nx_free")

    def upload(self, array):
        array = np.ascontiguousarray(array)
        dev = self.malloc(array.nbytes)
        check(self._h, lib().nx_memcpy_h2d(self._h, C.c_void_p(dev), _ptr(array), C.c_size_t(array.nbytes)), "h2d")
        return dev

    def download(self, dev, shape, dtype):
        out = np.empty(shape, dtype)
        check(self._h, lib().nx_memcpy_d2h(self._h, _ptr(out), C.c_void_p(dev), C.c_size_t(out.nbytes)), "d2h")
        return out


# ------------------------------------------------------------------------------------------------ builder ----
class BVH2:
    """NXB::BVH2 (vendor/NexusBVH/NexusBVH/include/NXB/BVH.h:18-37): device handle + ToHost."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def ToHost(self):
        """NXB::ToHost (BVHBuilder.h:40): (2n-1, 8) uint32 view of the 32-byte nodes."""
        nodes = np.empty((self.h.node_count, 8), np.uint32)
        check(self.ctx._h, lib().nx_bvh2_to_host(self.ctx._h, C.byref(self.h), _ptr(nodes)), "nx_bvh2_to_host")
        return nodes

    @property
    def bounds(self):
        return np.array(list(self.h.bounds.bmin) + list(self.h.bounds.bmax), np.float32)

    def Free(self):
        lib().nx_bvh2_free(self.ctx._h, C.byref(self.h))


class BVH8:
    """NXB::BVH8 (BVH.h:39-96)."""

    def __init__(self, ctx, handle, owned=True):
        self.ctx, self.h, self.owned = ctx, handle, owned

    @property
    def nodeCount(self):
        return self.h.node_count

    @property
    def primCount(self):
        return self.h.prim_count

    @property
    def bounds(self):
        return np.array(list(self.h.bounds.bmin) + list(self.h.bounds.bmax), np.float32)

    def ToHost(self):
        nodes = np.empty((self.h.node_count, 20), np.uint32)
        prim = np.empty(self.h.prim_count, np.uint32)
        check(self.ctx._h, lib().nx_bvh8_to_host(self.ctx._h, C.byref(self.h), _ptr(nodes), _ptr(prim)), "nx_bvh8_to_host")
        return nodes, prim

    def RefitAABB(self, bounds):
        """In-place refit of a BVH8 built over boxes: same topology and leaf order, node frames and child boxes recomputed bottom-up from
        `bounds` ((n, 6) host array, primitive order).  No counterpart in the reference (it rebuilds its TLAS); nx_bvh8_refit_aabb."""
        bounds = np.ascontiguousarray(bounds, np.float32).reshape(-1, 6)
        assert bounds.shape[0] == self.h.prim_count
        dev = self.ctx.upload(bounds)
        try:
            check(self.ctx._h, lib().nx_bvh8_refit_aabb(self.ctx._h, C.byref(self.h), C.c_void_p(dev)), "RefitBVH8")
        finally:
            self.ctx.free(dev)

    def Free(self):
        if self.owned:
            lib().nx_bvh8_free(self.ctx._h, C.byref(self.h))


def _prims_to_device(ctx, prims):
    prims = np.ascontiguousarray(prims, np.float32)
    if prims.ndim != 2 or prims.shape[1] not in (6, 9):
        raise ValueError("primitives must be (n, 9) triangles or (n, 6) AABBs")
    return prims, ctx.upload(prims), 1 if prims.shape[1] == 9 else 0


def BuildBVH2(ctx, prims, prioritizeSpeed=False, metrics=False):
    """NXB::BuildBVH2<PrimT> (BVHBuilder.h:19).  prims: host (n,9) triangles or (n,6) AABBs (uploaded here)."""
    prims, dev, tri = _prims_to_device(ctx, prims)
    cfg, m, out = BuildConfig(int(prioritizeSpeed)), BuildMetrics(), Bvh2()
    fn = lib().nx_bvh2_build_tri if tri else lib().nx_bvh2_build_aabb
    try:
        check(ctx._h, fn(ctx._h, C.c_void_p(dev), C.c_uint32(prims.shape[0]), C.byref(cfg), C.byref(m) if metrics else None, C.byref(out)), "BuildBVH2")
    finally:
        ctx.free(dev)
    bvh = BVH2(ctx, out)
    return (bvh, m.as_dict()) if metrics else bvh


COLLAPSE_REFERENCE_GPU, COLLAPSE_SAH_OPTIMAL = 0, 1


def BuildBVH8(ctx, prims, prioritizeSpeed=False, metrics=False, collapse=COLLAPSE_REFERENCE_GPU, maxLeafPrims=0):
    """NXB::BuildBVH8<PrimT> (BVHBuilder.h:31).  collapse: the reference GPU converter's rule (default, trees identical to
    NexusBVH's) or the SAH-optimal collapse of the reference's CPU BVH8Builder (BVH8Builder.cpp:31-199) on the GPU."""
    prims, dev, tri = _prims_to_device(ctx, prims)
    cfg, m, out = BuildConfig(int(prioritizeSpeed), int(collapse), int(maxLeafPrims)), BuildMetrics(), Bvh8()
    fn = lib().nx_bvh8_build_tri if tri else lib().nx_bvh8_build_aabb
    try:
        check(ctx._h, fn(ctx._h, C.c_void_p(dev), C.c_uint32(prims.shape[0]), C.byref(cfg), C.byref(m) if metrics else None, C.byref(out)), "BuildBVH8")
    finally:
        ctx.free(dev)
    bvh = BVH8(ctx, out)
    return (bvh, m.as_dict()) if metrics else bvh


def BuildBVH8Device(ctx, prims_dev, n, prim_type=1, prioritizeSpeed=False, metrics=False, collapse=COLLAPSE_REFERENCE_GPU, maxLeafPrims=0):
    """NXB::BuildBVH8<PrimT> on primitives already resident on the device (the reference's signature takes a device
    pointer, BVHBuilder.h:31).  Asynchronous on the context's stream unless metrics are requested."""
    cfg, m, out = BuildConfig(int(prioritizeSpeed), int(collapse), int(maxLeafPrims)), BuildMetrics(), Bvh8()
    fn = lib().nx_bvh8_build_tri if prim_type else lib().nx_bvh8_build_aabb
    check(ctx._h, fn(ctx._h, C.c_void_p(prims_dev), C.c_uint32(n), C.byref(cfg), C.byref(m) if metrics else None, C.byref(out)), "BuildBVH8")
    bvh = BVH8(ctx, out)
    return (bvh, m.as_dict()) if metrics else bvh


def BuildBLAS(ctx, triangles):
    """The BLAS Mesh::Mesh (N/Assets/Mesh.h:29-40) builds for a mesh, as a standalone unit of work: host (n, 9) triangles in, an owned
    BVH8 out, built with the context's scene settings (collapse mode, leaf size, Morton width) exactly as AssetManager.AddMesh
    would.  What a rank of a sharded scene build runs for the meshes it owns (nexus_b200.multigpu.build_scene_sharded)."""
    tris = np.ascontiguousarray(triangles, np.float32).reshape(-1, 9)
    out = Bvh8()
    check(ctx._h, lib().nx_scene_build_blas(ctx._h, _ptr(tris), C.c_uint32(tris.shape[0]), C.byref(out)), "BuildBLAS")
    return BVH8(ctx, out)


def BenchmarkBuild(ctx, prims_dev, n, prim_type, prioritizeSpeed, warmup, iters, collapse=COLLAPSE_REFERENCE_GPU, maxLeafPrims=0):
    """NXB::BenchmarkBuild (BVHBuildMetrics.h:63-108) on primitives already resident on the device."""
    cfg, m, nodes = BuildConfig(int(prioritizeSpeed), int(collapse), int(maxLeafPrims)), BuildMetrics(), C.c_uint32(0)
    check(ctx._h, lib().nx_bvh8_benchmark(ctx._h, C.c_void_p(prims_dev), C.c_uint32(n), C.c_int(prim_type), C.byref(cfg), C.c_int(warmup),
                                          C.c_int(iters), C.byref(m), C.byref(nodes)), "BenchmarkBuild")
    d = m.as_dict()
    d["node_count"] = nodes.value
    return d


def debug_morton(ctx, prims, bits64):
    """Parity hook: Morton codes (primitive order) exactly as the device computes them."""
    prims, dev, tri = _prims_to_device(ctx, prims)
    out = np.empty(prims.shape[0], np.uint64)
    try:
        check(ctx._h, lib().nx_bvh_debug_morton(ctx._h, C.c_void_p(dev), C.c_uint32(prims.shape[0]), C.c_int(tri), C.c_int(int(bits64)), _ptr(out)), "debug_morton")
    finally:
        ctx.free(dev)
    return out


# -------------------------------------------------------------------------------------------------- scene ----
class Material:
    """Material (src/Assets/Material.h:6-26), same defaults."""

    def __init__(self, **kw):
        self.baseColor = (0.8, 0.8, 0.8)
        self.metalness = 0.0
        self.roughness = 0.3
        self.anisotropy = 0.0
        self.specularWeight = 1.0
        self.specularColor = (1.0, 1.0, 1.0)
        self.ior = 1.5
        self.transmission = 0.0
        self.emissionColor = (1.0, 1.0, 1.0)
        self.intensity = 0.0
        self.opacity = 1.0
        self.baseColorMap = self.emissiveMap = self.normalMap = self.roughnessMap = self.metalnessMap = self.metallicRoughnessMap = -1
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)

    def pod(self):
        p = MaterialPod()
        p.base_color[:] = self.baseColor
        p.metalness, p.roughness, p.anisotropy, p.specular_weight = self.metalness, self.roughness, self.anisotropy, self.specularWeight
        p.specular_color[:] = self.specularColor
        p.ior, p.transmission = self.ior, self.transmission
        p.emission_color[:] = self.emissionColor
        p.intensity, p.opacity = self.intensity, self.opacity
        p.base_color_map, p.emissive_map, p.normal_map = self.baseColorMap, self.emissiveMap, self.normalMap
        p.roughness_map, p.metalness_map, p.metallic_roughness_map = self.roughnessMap, self.metalnessMap, self.metallicRoughnessMap
        return p


class Light:
    """Light (src/Scene/Light.h:10-54)."""
    POINT, SPOT, DIRECTIONAL, MESH = 0, 1, 2, 3

    def __init__(self, type, position=(0, 0, 0), direction=(0, -1, 0), color=(1, 1, 1), intensity=1.0, instance=0):
        self.type, self.position, self.direction, self.color, self.intensity, self.instance = type, position, direction, color, intensity, instance

    def pod(self):
        p = LightPod()
        p.type = self.type
        p.position[:] = self.position
        p.direction[:] = self.direction
        p.color[:] = self.color
        p.intensity = self.intensity
        p.instance = self.instance
        return p


class Camera:
    """Camera (src/Scene/Camera.h:9-52): position, forward, horizontal FOV (degrees), focus distance, defocus angle."""

    def __init__(self, position=(0.0, 4.0, 14.0), forward=(0.0, 0.0, -1.0), horizontalFOV=45.0, focusDistance=5.0, defocusAngle=0.0, right=(0.0, 0.0, 0.0)):
        self.position, self.forward, self.right = position, forward, right
        self.horizontalFOV, self.focusDistance, self.defocusAngle = horizontalFOV, focusDistance, defocusAngle
        self._invalid = True

    # Camera.h:19-35: the setters do not invalidate by themselves in the reference either; Scene::Update pushes an invalid camera
    def GetHorizontalFOV(self): return self.horizontalFOV
    def SetHorizontalFOV(self, v): self.horizontalFOV = float(v)
    def GetDefocusAngle(self): return self.defocusAngle
    def SetDefocusAngle(self, v): self.defocusAngle = float(v)
    def GetFocusDist(self): return self.focusDistance
    def SetFocusDist(self, v): self.focusDistance = float(v)
    def GetPosition(self): return self.position
    def SetPosition(self, v): self.position = tuple(float(x) for x in v)
    def GetForwardDirection(self): return self.forward
    def SetForwardDirection(self, v): self.forward = tuple(float(x) for x in v)
    def GetRightDirection(self): return self.right
    def SetRightDirection(self, v): self.right = tuple(float(x) for x in v)
    def IsInvalid(self): return self._invalid
    def SetInvalid(self, invalid): self._invalid = bool(invalid)
    def Invalidate(self): self._invalid = True

    def pod(self):
        p = CameraPod()
        p.position[:] = self.position
        p.forward[:] = self.forward
        p.right[:] = self.right
        p.horizontal_fov_deg, p.focus_distance, p.defocus_angle_deg = self.horizontalFOV, self.focusDistance, self.defocusAngle
        return p


class RenderSettings:
    """RenderSettings (src/Renderer/RenderSettings.h:5-17)."""

    def __init__(self, useMIS=True, pathLength=10, backgroundColor=(0.0, 0.0, 0.0), backgroundIntensity=1.0, toneMapping=3, exposure=0.0):
        self.useMIS, self.pathLength, self.backgroundColor = useMIS, pathLength, backgroundColor
        self.backgroundIntensity, self.toneMapping, self.exposure = backgroundIntensity, toneMapping, exposure

    def pod(self):
        p = RenderSettingsPod()
        p.use_mis, p.path_length = int(self.useMIS), int(self.pathLength)
        p.background_color[:] = self.backgroundColor
        p.background_intensity, p.tone_mapping, p.exposure = self.backgroundIntensity, self.toneMapping, self.exposure
        return p


class MeshInstance:
    """MeshInstance (src/Scene/MeshInstance.h:9-77): host object with position / rotation (Euler degrees) / scale, mesh and material
    index.  As in the reference, edits take effect at the next Scene.Update() after Scene.InvalidateMeshInstance(index); the
    setters here also invalidate by themselves, so forgetting the call is harmless.  (The reference's SetRotationY / SetRotationZ
    write `position`, MeshInstance.h:24-25; that bug is not reproduced.)"""

    def __init__(self, scene, index, meshIdx, materialIdx=-1, position=(0.0, 0.0, 0.0), rotation=(0.0, 0.0, 0.0), scale=(1.0, 1.0, 1.0), matrix=None):
        self.scene, self.index, self.meshIdx, self.materialIdx = scene, int(index), int(meshIdx), int(materialIdx)
        self.position, self.rotation, self.scale = tuple(map(float, position)), tuple(map(float, rotation)), tuple(map(float, scale))
        self._matrix = None if matrix is None else np.array(matrix, np.float32).reshape(4, 4)    # explicit-matrix instances (asset import)
        self.name = f"instance {index}"
        self._material_edited = self._trs_edited = False

    def _edited(self):
        self._matrix = None
        self._trs_edited = True
        self.scene.InvalidateMeshInstance(self.index)

    def SetPosition(self, p):
        self.position = tuple(map(float, p)); self._edited()

    def SetRotationX(self, r):
        self.rotation = (float(r), self.rotation[1], self.rotation[2]); self._edited()

    def SetRotationY(self, r):
        self.rotation = (self.rotation[0], float(r), self.rotation[2]); self._edited()

    def SetRotationZ(self, r):
        self.rotation = (self.rotation[0], self.rotation[1], float(r)); self._edited()

    def SetScale(self, s):
        self.scale = (float(s),) * 3 if np.isscalar(s) else tuple(map(float, s)); self._edited()

    def SetTransform(self, position, rotation, scale):
        self.position, self.rotation, self.scale = tuple(map(float, position)), tuple(map(float, rotation)), tuple(map(float, scale))
        self._edited()
        self.scene._flush_instance(self)        # applied at once (older callers trace right after SetTransform without Update)

    def AssignMaterial(self, materialIdx):
        self.materialIdx = int(materialIdx)
        self._material_edited = True
        self.scene.InvalidateMeshInstance(self.index)

    def GetTransfromationMatrix(self):
        """T * Rz * Ry * Rx * S (MeshInstance.h:36-40; the spelling is the reference's), row-major 4x4, as the device sees it."""
        self.scene._flush_instance(self)
        m = np.zeros(16, np.float32)
        check(self.scene.ctx._h, lib().nx_scene_instance_matrix(self.scene._h, C.c_uint32(self.index), _ptr(m)), "GetTransfromationMatrix")
        return m.reshape(4, 4)

    def GetBounds(self):
        """World AABB of the eight transformed corners of the mesh's box (MeshInstance.h:42-53): (bmin, bmax) as 6 floats."""
        self.scene._flush_instance(self)
        b = Aabb()
        check(self.scene.ctx._h, lib().nx_scene_instance_bounds(self.scene._h, C.c_uint32(self.index), C.byref(b)), "GetBounds")
        return np.array(list(b.bmin) + list(b.bmax), np.float32)


class AssetManager:
    """AssetManager (src/Assets/AssetManager.h:13-57): meshes and materials of a scene."""

    def __init__(self, scene):
        self.scene = scene
        self._materials, self._invalid_materials = [], set()
        self._texture_count = 0

    def AddMaterial(self, material):
        p = material.pod()
        idx = check(self.scene.ctx._h, lib().nx_scene_add_material(self.scene._h, C.byref(p)), "AddMaterial")
        self._materials.append(material)
        return idx

    def GetMaterials(self):
        """The host materials (AssetManager.h:24); edit one, then InvalidateMaterial(index)."""
        return self._materials

    def SendDataToDevice(self):
        """AssetManager::SendDataToDevice (AssetManager.cpp:74-84): pushes the invalidated materials; True if there were any."""
        dirty = sorted(self._invalid_materials)
        for i in dirty:
            p = self._materials[i].pod()
            check(self.scene.ctx._h, lib().nx_scene_set_material(self.scene._h, C.c_uint32(i), C.byref(p)), "InvalidateMaterial")
        self._invalid_materials.clear()
        return bool(dirty)

    def IsInvalid(self):
        return bool(self._invalid_materials)

    def AddMesh(self, name, materialIdx, triangles, triangleData=None):
        """AddMesh(name, materialIdx, triangles, triangleData) (AssetManager.cpp:24-33): builds the BLAS immediately."""
        tris = np.ascontiguousarray(triangles, np.float32).reshape(-1, 9)
        td = None if triangleData is None else np.ascontiguousarray(triangleData, np.float32).reshape(-1, 24)
        if td is not None and td.shape[0] != tris.shape[0]:
            raise ValueError("triangleData must have one row per triangle")
        return check(self.scene.ctx._h, lib().nx_scene_add_mesh(self.scene._h, _ptr(tris), _ptr(td) if td is not None else None,
                                                                 C.c_uint32(tris.shape[0]), C.c_uint32(materialIdx)), f"AddMesh({name})")

    def AddMeshPrebuilt(self, name, materialIdx, triangles, triangleData, nodes_dev, node_count, prim_idx_dev, bounds):
        """AddMesh with the BLAS supplied (DEVICE pointers to node_count 80-byte nodes and one uint32 per triangle, e.g. inside an
        all-gather buffer; copied) instead of built: the receiving side of a sharded scene build."""
        tris = np.ascontiguousarray(triangles, np.float32).reshape(-1, 9)
        td = None if triangleData is None else np.ascontiguousarray(triangleData, np.float32).reshape(-1, 24)
        if td is not None and td.shape[0] != tris.shape[0]:
            raise ValueError("triangleData must have one row per triangle")
        b = np.asarray(bounds, np.float32).reshape(6)
        box = Aabb((C.c_float * 3)(*b[:3]), (C.c_float * 3)(*b[3:]))
        return check(self.scene.ctx._h, lib().nx_scene_add_mesh_prebuilt(self.scene._h, _ptr(tris), _ptr(td) if td is not None else None,
                                                                          C.c_uint32(tris.shape[0]), C.c_uint32(materialIdx), C.c_void_p(int(nodes_dev)),
                                                                          C.c_uint32(int(node_count)), C.c_void_p(int(prim_idx_dev)), C.byref(box)), f"AddMeshPrebuilt({name})")

    def AddTexture(self, pixels, sRGB=False):
        """AddTexture (AssetManager.h:31) + Texture::ToDevice (Texture.cpp:12-46).  pixels: (h, w, 4) uint8 (normalised reads,
        sRGB decoded by the sampler when sRGB) or float32 (HDR).  Returns the index Material.*Map refers to."""
        pixels = np.ascontiguousarray(pixels)
        if pixels.ndim != 3 or pixels.shape[2] != 4 or pixels.dtype not in (np.uint8, np.float32):
            raise ValueError("texture pixels must be (h, w, 4) uint8 or float32")
        idx = check(self.scene.ctx._h, lib().nx_scene_add_texture(self.scene._h, _ptr(pixels), C.c_uint32(pixels.shape[1]), C.c_uint32(pixels.shape[0]),
                                                                   C.c_int(int(pixels.dtype == np.float32)), C.c_int(int(sRGB))), "AddTexture")
        self._texture_count = idx + 1
        return idx

    def InvalidateMaterial(self, index, material=None):
        """AssetManager::InvalidateMaterial (AssetManager.h:36): marks GetMaterials()[index] for upload at the next Scene.Update().
        With `material` given it replaces the entry and uploads at once."""
        if not (0 <= index < len(self._materials)):
            raise NexusError(f"InvalidateMaterial: material {index} of {len(self._materials)}")
        if material is not None:
            self._materials[index] = material
        self._invalid_materials.add(int(index))
        if material is not None:
            self.SendDataToDevice()


class Scene:
    """Scene (src/Scene/Scene.h:16-77)."""

    def __init__(self, ctx, resolution):
        self.ctx = ctx
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self._h = C.c_void_p()
        check(ctx._h, lib().nx_scene_create(ctx._h, C.c_uint32(self.resolution[0]), C.c_uint32(self.resolution[1]), C.byref(self._h)), "Scene")
        self._assets = AssetManager(self)
        self._camera, self._settings = Camera(), RenderSettings()       # Scene::Scene defaults (Scene.cpp:8-12, RenderSettings.h:5-17)
        self._instances, self._invalid_instances = [], set()
        self._lights, self._invalid_lights = [], set()

    def close(self):
        if self._h:
            lib().nx_scene_destroy(self._h)
            self._h = C.c_void_p()

    def GetAssetManager(self):
        return self._assets

    def OnResize(self, resolution):
        """Camera::OnResize (src/Scene/Camera.cpp:118-128): the output resolution changed; only the camera record depends on it.
        Call PathTracer.OnResize with the same resolution."""
        self.resolution = (int(resolution[0]), int(resolution[1]))
        check(self.ctx._h, lib().nx_scene_set_resolution(self._h, C.c_uint32(self.resolution[0]), C.c_uint32(self.resolution[1])), "Scene.OnResize")

    def AddMaterial(self, material):
        return self._assets.AddMaterial(material)

    def GetMaterials(self):
        return self._assets.GetMaterials()

    def CreateMeshInstance(self, meshId, materialIdx=-1, position=(0, 0, 0), rotation=(0, 0, 0), scale=(1, 1, 1)):
        p, r, s = (np.asarray(v, np.float32) for v in (position, rotation, scale))
        idx = check(self.ctx._h, lib().nx_scene_add_instance(self._h, C.c_uint32(meshId), C.c_int32(materialIdx), _ptr(p), _ptr(r), _ptr(s)), "CreateMeshInstance")
        inst = MeshInstance(self, idx, meshId, materialIdx, position, rotation, scale)
        self._instances.append(inst)
        return inst

    def CreateMeshInstanceMatrix(self, meshId, matrix, materialIdx=-1):
        m = np.ascontiguousarray(matrix, np.float32).reshape(16)
        idx = check(self.ctx._h, lib().nx_scene_add_instance_matrix(self._h, C.c_uint32(meshId), C.c_int32(materialIdx), _ptr(m)), "CreateMeshInstance")
        inst = MeshInstance(self, idx, meshId, materialIdx, matrix=m)
        self._instances.append(inst)
        return inst

    def GetMeshInstances(self):
        return self._instances

    def CreateMeshInstanceFromFile(self, filePath, fileName="", decode_image=None):
        """Scene::CreateMeshInstanceFromFile (Scene.cpp:97-100 -> OBJLoader::LoadOBJ, OBJLoader.cpp:420-446): imports an asset
        into THIS scene - its materials and textures are appended to the asset manager, every primitive becomes a mesh and every
        node that references it an instance.  .glb through nexus_b200.gltf, .obj through nexus_b200.obj (the reference goes
        through Assimp).  Returns the new MeshInstance objects; the asset's camera, if any, is ignored as in the reference's
        loader when a scene already exists."""
        import os as _os
        path = _os.fspath(filePath) + fileName
        ext = _os.path.splitext(path)[1].lower()
        if ext == ".glb":
            from .gltf import load_glb
            d = load_glb(path, decode_image=decode_image)
        elif ext == ".obj":
            from .obj import load_obj
            d = load_obj(path)
        else:
            raise NexusError(f"CreateMeshInstanceFromFile: unsupported asset type '{ext}' (.glb and .obj are)")
        am = self._assets
        tex0 = am._texture_count
        for pixels, srgb in d.get("textures", []):
            am.AddTexture(pixels, srgb)
        mat0 = len(am._materials)
        for m in d["materials"]:
            for attr in ("baseColorMap", "emissiveMap", "normalMap", "roughnessMap", "metalnessMap", "metallicRoughnessMap"):
                if getattr(m, attr) >= 0:
                    setattr(m, attr, getattr(m, attr) + tex0)
            am.AddMaterial(m)
        mesh_ids = [am.AddMesh(m["name"], mat0 + m["material"], m["triangles"], m.get("triangle_data")) for m in d["meshes"]]
        created = []
        for i in d["instances"]:
            mat = i.get("material", -1)
            inst = self.CreateMeshInstanceMatrix(mesh_ids[i["mesh"]], i["matrix"], mat0 + mat if mat >= 0 else -1)
            inst.name = d["meshes"][i["mesh"]]["name"]
            created.append(inst)
        return created

    def IsEmpty(self):
        return not self._instances

    def InvalidateMeshInstance(self, instanceId):
        """Scene::InvalidateMeshInstance (Scene.cpp:109-112): the instance's transform / material are re-read at the next Update()."""
        self._invalid_instances.add(int(instanceId))

    def _flush_instance(self, inst):
        if inst.index not in self._invalid_instances:
            return
        if inst._matrix is None and inst._trs_edited:        # position / rotation / scale were set: T * Rz * Ry * Rx * S, composed by the library
            p, r, s = (np.asarray(v, np.float32) for v in (inst.position, inst.rotation, inst.scale))
            check(self.ctx._h, lib().nx_scene_set_instance_transform(self._h, C.c_uint32(inst.index), _ptr(p), _ptr(r), _ptr(s)), "InvalidateMeshInstance")
            inst._trs_edited = False
        if inst._material_edited:
            check(self.ctx._h, lib().nx_scene_set_instance_material(self._h, C.c_uint32(inst.index), C.c_int32(inst.materialIdx)), "AssignMaterial")
            inst._material_edited = False
        self._invalid_instances.discard(inst.index)

    def AddLight(self, light):
        p = light.pod()
        idx = check(self.ctx._h, lib().nx_scene_add_light(self._h, C.byref(p)), "AddLight")
        self._lights.append(light)
        return idx

    def GetLights(self):
        """The lights added through AddLight (Scene.h:46); edit one, then InvalidateLight(index).  Emissive instances become
        lights by themselves (Scene.cpp:157-219) and are not in this list."""
        return self._lights

    def InvalidateLight(self, lightIdx):
        if not (0 <= lightIdx < len(self._lights)):
            raise NexusError(f"InvalidateLight: light {lightIdx} of {len(self._lights)}")
        self._invalid_lights.add(int(lightIdx))

    def RemoveLight(self, index):
        """Scene::RemoveLight (Scene.cpp:129-132): later lights move down by one."""
        if not (0 <= index < len(self._lights)):
            raise NexusError(f"RemoveLight: light {index} of {len(self._lights)}")
        check(self.ctx._h, lib().nx_scene_remove_light(self._h, C.c_uint32(index)), "RemoveLight")
        del self._lights[index]
        self._invalid_lights = {i - (i > index) for i in self._invalid_lights if i != index}

    def GetCamera(self):
        """The host camera (Scene.h:22); edit it, call Invalidate() on it, and the next Update() uploads it."""
        return self._camera

    def SetCamera(self, camera):
        p = camera.pod()
        check(self.ctx._h, lib().nx_scene_set_camera(self._h, C.byref(p)), "SetCamera")
        self._camera = camera
        camera.SetInvalid(False)

    def GetRenderSettings(self):
        """The host render settings (Scene.h:27-28); edits are uploaded by the next Update()."""
        return self._settings

    def SetRenderSettings(self, rs):
        p = rs.pod()
        check(self.ctx._h, lib().nx_scene_set_render_settings(self._h, C.byref(p)), "SetRenderSettings")
        self._settings = rs
        self._settings_sent = (rs.useMIS, rs.pathLength, tuple(rs.backgroundColor), rs.backgroundIntensity, rs.toneMapping, rs.exposure)

    def IsInvalid(self):
        """Scene::IsInvalid (Scene.h:32): something was edited since the last Update()."""
        rs = self._settings
        return bool(self._invalid_instances or self._invalid_lights or self._camera.IsInvalid() or self._assets.IsInvalid()
                    or getattr(self, "_settings_sent", None) != (rs.useMIS, rs.pathLength, tuple(rs.backgroundColor), rs.backgroundIntensity, rs.toneMapping, rs.exposure))

    def AddHDRMap(self, rgba, fileName=None):
        """Scene::AddHDRMap (Scene.cpp:102-107).  Either pixels, (h, w, 4) float32 equirect, or the reference's
        (filePath, fileName) pair naming a Radiance .hdr file (read by nexus_b200.hdr.load_hdr as stb_image would)."""
        if isinstance(rgba, (str, bytes)) or hasattr(rgba, "__fspath__"):
            import os as _os
            from .hdr import load_hdr
            rgba = load_hdr(_os.fspath(rgba) + (fileName or ""))
        rgba = np.ascontiguousarray(rgba, np.float32)
        check(self.ctx._h, lib().nx_scene_set_hdr_map(self._h, _ptr(rgba), C.c_uint32(rgba.shape[1]), C.c_uint32(rgba.shape[0])), "AddHDRMap")

    def Update(self):
        """Scene::Update (Scene.cpp:34-63): uploads what was invalidated (camera, render settings, materials, instances, lights),
        rebuilds the TLAS when an instance changed and refreshes the light list."""
        if self._camera.IsInvalid():
            self.SetCamera(self._camera)
        rs = self._settings
        if getattr(self, "_settings_sent", None) != (rs.useMIS, rs.pathLength, tuple(rs.backgroundColor), rs.backgroundIntensity, rs.toneMapping, rs.exposure):
            self.SetRenderSettings(rs)
        self._assets.SendDataToDevice()
        for i in sorted(self._invalid_instances):
            self._flush_instance(self._instances[i]) if i < len(self._instances) else self._invalid_instances.discard(i)
        for i in sorted(self._invalid_lights):
            p = self._lights[i].pod()
            check(self.ctx._h, lib().nx_scene_set_light(self._h, C.c_uint32(i), C.byref(p)), "InvalidateLight")
        self._invalid_lights.clear()
        check(self.ctx._h, lib().nx_scene_update(self._h), "Scene::Update")

    def BuildTLAS(self):
        """Scene::BuildTLAS (Scene.cpp:65-78).  The library rebuilds the TLAS inside Update whenever an instance changed; this is
        Update under the reference's name for callers that drive the two steps themselves."""
        self.Update()

    def MeshBounds(self, meshId):
        b = Aabb()
        check(self.ctx._h, lib().nx_scene_mesh_bounds(self._h, C.c_uint32(meshId), C.byref(b)), "MeshBounds")
        return np.array(list(b.bmin) + list(b.bmax), np.float32)

    def MeshBVH(self, meshId):
        h = Bvh8()
        check(self.ctx._h, lib().nx_scene_mesh_bvh(self._h, C.c_uint32(meshId), C.byref(h)), "MeshBVH")
        return BVH8(self.ctx, h, owned=False)

    def TLAS(self):
        h = Bvh8()
        check(self.ctx._h, lib().nx_scene_tlas(self._h, C.byref(h)), "TLAS")
        return BVH8(self.ctx, h, owned=False)

    def TlasHistory(self):
        """(builds, refits): how often this scene's TLAS was built from scratch / refitted in place."""
        b, r = C.c_uint32(0), C.c_uint32(0)
        check(self.ctx._h, lib().nx_scene_tlas_history(self._h, C.byref(b), C.byref(r)), "TlasHistory")
        return b.value, r.value

    def ExportTlasEntries(self):
        """The TLAS is built over entries: the instances that keep a BLAS of their own, then the merged BLAS.  Returns the instance id
        of every entry (0xffffffff for the merged BLAS); TLAS().ToHost()'s primitive indices are entry numbers."""
        n = C.c_uint32(0)
        check(self.ctx._h, lib().nx_scene_export_tlas_entries(self._h, None, C.byref(n)), "ExportTlasEntries")
        out = np.zeros(max(n.value, 1), np.uint32)
        check(self.ctx._h, lib().nx_scene_export_tlas_entries(self._h, _ptr(out), C.byref(n)), "ExportTlasEntries")
        return out[:n.value]

    def ExportMerged(self, bounds=True):
        """The merged BLAS (Context.SetInstanceMerging; world-space nodes, object-space triangle tests), or None:
        dict(bvh, bounds (n, 6) padded world boxes the tree was built over, instance (n,), prim (n,))."""
        n, h = C.c_uint32(0), Bvh8()
        check(self.ctx._h, lib().nx_scene_export_merged(self._h, C.byref(h), None, None, None, C.byref(n)), "ExportMerged")
        if not n.value:
            return None
        box = np.empty((n.value, 6), np.float32) if bounds else None
        inst, prim = np.empty(n.value, np.uint32), np.empty(n.value, np.uint32)
        check(self.ctx._h, lib().nx_scene_export_merged(self._h, C.byref(h), _ptr(box) if bounds else None, _ptr(inst), _ptr(prim), C.byref(n)), "ExportMerged")
        return {"bvh": BVH8(self.ctx, h, owned=False), "bounds": box, "instance": inst, "prim": prim}

    # reference device layouts, for the reference arm of parity tests
    def ExportInstances(self):
        n = C.c_uint32(0)
        check(self.ctx._h, lib().nx_scene_export_instances(self._h, None, C.byref(n)), "ExportInstances")
        out = np.zeros((n.value, 160), np.uint8)
        check(self.ctx._h, lib().nx_scene_export_instances(self._h, _ptr(out), C.byref(n)), "ExportInstances")
        return out

    def ExportCamera(self):
        out = np.zeros(88, np.uint8)
        check(self.ctx._h, lib().nx_scene_export_camera(self._h, _ptr(out)), "ExportCamera")
        return out

    def ExportLights(self):
        n = C.c_uint32(0)
        check(self.ctx._h, lib().nx_scene_export_lights(self._h, None, C.byref(n)), "ExportLights")
        out = np.zeros((max(n.value, 1), 52), np.uint8)
        check(self.ctx._h, lib().nx_scene_export_lights(self._h, _ptr(out), C.byref(n)), "ExportLights")
        return out[:n.value]

    # traversal parity hooks
    def TraceClosest(self, rays, timed=False):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        ms = C.c_float(0)
        check(self.ctx._h, lib().nx_trace_closest(self._h, _ptr(rays), C.c_uint32(rays.shape[0]), _ptr(hits), C.byref(ms) if timed else None), "TraceClosest")
        return (hits, ms.value) if timed else hits

    def TraceAny(self, rays):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        occ = np.empty(rays.shape[0], np.uint8)
        check(self.ctx._h, lib().nx_trace_any(self._h, _ptr(rays), C.c_uint32(rays.shape[0]), _ptr(occ), None), "TraceAny")
        return occ

    def TraceClosestDevice(self, rays_dev, n, hits_dev):
        ms = C.c_float(0)
        check(self.ctx._h, lib().nx_trace_closest_device(self._h, C.c_void_p(rays_dev), C.c_uint32(n), C.c_void_p(hits_dev), C.byref(ms)), "TraceClosestDevice")
        return ms.value

    def TraceStats(self, rays_dev, n, hits_dev):
        out = (C.c_uint64 * 4)()
        check(self.ctx._h, lib().nx_trace_stats(self._h, C.c_void_p(rays_dev), C.c_uint32(n), C.c_void_p(hits_dev), out), "TraceStats")
        return {"nodes": out[0], "tris": out[1], "insts": out[2], "rays": out[3]}


def make_rays(origins, directions, tmax=1.0e30):
    rays = np.zeros(len(origins), RAY_DTYPE)
    rays["origin"] = origins
    rays["direction"] = directions
    rays["tmax"] = tmax
    return rays


# ----------------------------------------------------------------------------------------------- renderer ----
class PathTracer:
    """PathTracer (src/Renderer/PathTracer.h:9-66): owns the wavefront queues and the accumulation buffer."""

    def __init__(self, ctx, resolution):
        self.ctx = ctx
        self.resolution = (int(resolution[0]), int(resolution[1]))
        self._h = C.c_void_p()
        check(ctx._h, lib().nx_renderer_create(ctx._h, C.c_uint32(self.resolution[0]), C.c_uint32(self.resolution[1]), C.byref(self._h)), "PathTracer")
        self._frame = 0

    def close(self):
        if self._h:
            lib().nx_renderer_destroy(self._h)
            self._h = C.c_void_p()

    def ResetFrameNumber(self):
        check(self.ctx._h, lib().nx_renderer_reset_accumulation(self._h), "ResetFrameNumber")
        self._frame = 0

    def OnResize(self, resolution):
        self.resolution = (int(resolution[0]), int(resolution[1]))
        check(self.ctx._h, lib().nx_renderer_resize(self._h, C.c_uint32(self.resolution[0]), C.c_uint32(self.resolution[1])), "OnResize")
        self._frame = 0

    def Render(self, scene, frames=1, firstFrame=None):
        """PathTracer::Render (PathTracer.cpp:166-200): one call renders `frames` consecutive frame numbers (1 spp each)."""
        first = self._frame + 1 if firstFrame is None else int(firstFrame)
        check(self.ctx._h, lib().nx_renderer_render(self._h, scene._h, C.c_uint32(first), C.c_uint32(frames)), "Render")
        self._frame = first + frames - 1

    def GetFrameNumber(self):
        return lib().nx_renderer_frame_count(self._h)

    def Reset(self):
        """PathTracer::Reset (PathTracer.cpp:61-159) re-allocates the queues for the current resolution and clears the
        accumulation; the queues here are sized once per resolution, so this is ResetFrameNumber."""
        self.ResetFrameNumber()

    def UpdateDeviceScene(self, scene):
        """PathTracer::UpdateDeviceScene (PathTracer.cpp:216-219) copies the D_Scene into the device symbol; here the scene view is
        a kernel parameter block assembled per Render call, so all that is left is flushing the scene's pending edits."""
        scene.Update()

    def Stats(self):
        st = FrameStats()
        check(self.ctx._h, lib().nx_renderer_stats(self._h, C.byref(st)), "Stats")
        return {n: getattr(st, n) for n, _ in st._fields_}

    KERNELS = ("generate", "trace_closest", "shade", "trace_any")

    def SetProfiling(self, events=True, work=False):
        """Measurement hook: per-kernel CUDA-event times (events) and traversal work counters (work) for the next Render calls."""
        check(self.ctx._h, lib().nx_renderer_set_profiling(self._h, C.c_int(int(events) | (int(work) << 1))), "SetProfiling")

    def Profile(self):
        p = KernelProfile()
        check(self.ctx._h, lib().nx_renderer_profile(self._h, C.byref(p)), "Profile")
        out = {k: {"ms": p.ms[i], "launches": p.launches[i]} for i, k in enumerate(self.KERNELS)}
        for name, arr in (("closest_work", p.closest_work), ("any_work", p.any_work)):
            out[name] = {"nodes": arr[0], "tris": arr[1], "insts": arr[2], "rays": arr[3]}
        for name, arr in (("closest_sched", p.closest_sched), ("any_sched", p.any_sched)):
            out[name] = dict(zip(("iters", "lanes_node", "tri_rounds", "tri_lanes", "setup_rounds", "setup_lanes", "sphere_culled", "node_rounds", "fetch_rounds", "fetch_lanes"), list(arr)))
        return out

    def ReadAccumulation(self, out=None):
        """Linear radiance mean, (h, w, 3) float32, row 0 = bottom row (the reference's pixel order)."""
        w, h = self.resolution
        if out is None:
            out = np.empty((h, w, 3), np.float32)
        check(self.ctx._h, lib().nx_renderer_read_accum(self._h, _ptr(out)), "ReadAccumulation")
        return out

    def AccumulationDevice(self):
        p, n = C.c_void_p(), C.c_uint32(0)
        check(self.ctx._h, lib().nx_renderer_accum_device(self._h, C.byref(p), C.byref(n)), "AccumulationDevice")
        return p.value, n.value

    def SetAccumulatedFrames(self, frames):
        check(self.ctx._h, lib().nx_renderer_set_accum_frames(self._h, C.c_uint32(frames)), "SetAccumulatedFrames")

    def ReadRGBA8(self, scene, out=None):
        """Display transform of AccumulateKernel (PathTracer.cu:527-548) on the running mean, (h, w) packed RGBA8."""
        w, h = self.resolution
        if out is None:
            out = np.empty((h, w), np.uint32)
        check(self.ctx._h, lib().nx_renderer_read_rgba8(self._h, scene._h, _ptr(out)), "ReadRGBA8")
        return out

    def Present(self, scene, out):
        """Pipelined display read-back (the reference's PBO path, PathTracer.cpp:170-199 + Renderer.cpp:41-48): queues the
        display transform of the current accumulation and its copy into `out` ((h, w) uint32, pinned host memory for a truly
        asynchronous copy) behind the frames already submitted and returns a ticket at once."""
        w, h = self.resolution
        assert out.dtype == np.uint32 and out.size == w * h and out.flags["C_CONTIGUOUS"]
        t = C.c_int(0)
        check(self.ctx._h, lib().nx_renderer_present(self._h, scene._h, _ptr(out), C.byref(t)), "Present")
        return t.value

    def PresentDevice(self, scene, dev_ptr):
        """The display transform of the current accumulation written straight into caller-supplied device memory (W * H RGBA8 words): what
        the reference's OpenGL viewer does with its mapped pixel buffer (PixelBuffer.cpp:4-40).  Queued on the context's stream."""
        check(self.ctx._h, lib().nx_renderer_present_device(self._h, scene._h, C.c_void_p(int(dev_ptr))), "PresentDevice")

    def PresentWait(self, ticket):
        """Blocks until the image of `ticket` is in host memory; returns the queue totals of the render call it shows."""
        st = FrameStats()
        check(self.ctx._h, lib().nx_renderer_present_wait(self._h, C.c_int(ticket), C.byref(st)), "PresentWait")
        return {n: getattr(st, n) for n, _ in st._fields_}

    def SetPixelQuery(self, x, y):
        """PathTracer::SetPixelQuery (PathTracer.cpp:233-240): asks for the instance under pixel (x, y), row 0 = bottom row."""
        check(self.ctx._h, lib().nx_renderer_set_pixel_query(self._h, C.c_uint32(x), C.c_uint32(y)), "SetPixelQuery")

    def PixelQueryPending(self):
        return lib().nx_renderer_pixel_query_pending(self._h) == 1

    def SynchronizePixelQuery(self):
        """PathTracer::SynchronizePixelQuery (PathTracer.cpp:221-231): instance id seen by the queried pixel's primary ray in
        the frame rendered after SetPixelQuery, -1 for a miss."""
        v = C.c_int32(-1)
        check(self.ctx._h, lib().nx_renderer_sync_pixel_query(self._h, C.byref(v)), "SynchronizePixelQuery")
        self._selected = v.value
        return v.value

    def GetSelectedInstance(self):
        return getattr(self, "_selected", -1)


TONE_NONE, TONE_ACES, TONE_UNCHARTED2, TONE_AGX_DEFAULT, TONE_AGX_GOLDEN, TONE_AGX_PUNCHY = range(6)   # ColorUtils::ToneMapping


def display_transform(ctx, rgb, toneMapping=TONE_AGX_DEFAULT, exposure=0.0):
    """AccumulateKernel's display transform (PathTracer.cu:527-548, ColorUtils.h:27-212) of a linear float RGB image -> packed RGBA8."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.empty(rgb.shape[:-1], np.uint32)
    check(ctx._h, lib().nx_display_transform(ctx._h, _ptr(rgb), C.c_uint32(out.size), C.c_int(int(toneMapping)), C.c_float(exposure), _ptr(out)), "display_transform")
    return out


def write_pfm(path, rgb):
    """rgb: (h, w, 3) float32 in the renderer's layout, row 0 = bottom row (what ReadAccumulation returns); the file is upright."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    rc = lib().nx_write_pfm(str(path).encode(), _ptr(rgb), C.c_uint32(rgb.shape[1]), C.c_uint32(rgb.shape[0]))
    if rc < 0:
        raise NexusError(f"nx_write_pfm({path}) failed")


def write_exr(path, rgb):
    """rgb: (h, w, 3) float32 in the renderer's layout, row 0 = bottom row (what ReadAccumulation returns); the file is upright."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    rc = lib().nx_write_exr(str(path).encode(), _ptr(rgb), C.c_uint32(rgb.shape[1]), C.c_uint32(rgb.shape[0]))
    if rc < 0:
        raise NexusError(f"nx_write_exr({path}) failed")
