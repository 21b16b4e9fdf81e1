"""ctypes declarations for libnexus_b200.so (see include/nexus_b200.h).  No CPU fallback: importing works anywhere,
creating a context without the CUDA library or without a GPU raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NEXUS_B200_LIB") or os.path.join(_HERE, "libnexus_b200.so")   # override: kernel-variant experiments only


class NexusError(RuntimeError):
    pass


class Aabb(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3)]


class Bvh2(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("node_count", C.c_uint32), ("prim_count", C.c_uint32), ("bounds", Aabb)]


class Bvh8(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("node_count", C.c_uint32), ("prim_idx", C.c_void_p), ("prim_count", C.c_uint32), ("bounds", Aabb)]


class BuildConfig(C.Structure):
    _fields_ = [("prioritize_speed", C.c_int), ("collapse", C.c_int), ("max_leaf_prims", C.c_int)]


class BuildMetrics(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("scene_bounds_ms", "morton_ms", "sort_ms", "bvh2_ms", "bvh8_ms", "total_ms",
                                         "bvh2_cost", "bvh8_cost", "avg_children_per_node")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class MaterialPod(C.Structure):
    _fields_ = [("base_color", C.c_float * 3), ("metalness", C.c_float), ("roughness", C.c_float), ("anisotropy", C.c_float),
                ("specular_weight", C.c_float), ("specular_color", C.c_float * 3), ("ior", C.c_float), ("transmission", C.c_float),
                ("emission_color", C.c_float * 3), ("intensity", C.c_float), ("opacity", C.c_float),
                ("base_color_map", C.c_int32), ("emissive_map", C.c_int32), ("normal_map", C.c_int32), ("roughness_map", C.c_int32),
                ("metalness_map", C.c_int32), ("metallic_roughness_map", C.c_int32)]


class LightPod(C.Structure):
    _fields_ = [("type", C.c_int32), ("position", C.c_float * 3), ("direction", C.c_float * 3), ("color", C.c_float * 3),
                ("intensity", C.c_float), ("falloff_start", C.c_float), ("falloff_end", C.c_float), ("instance", C.c_uint32)]


class CameraPod(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("forward", C.c_float * 3), ("right", C.c_float * 3),
                ("horizontal_fov_deg", C.c_float), ("focus_distance", C.c_float), ("defocus_angle_deg", C.c_float)]


class RenderSettingsPod(C.Structure):
    _fields_ = [("use_mis", C.c_int32), ("path_length", C.c_int32), ("background_color", C.c_float * 3),
                ("background_intensity", C.c_float), ("tone_mapping", C.c_int32), ("exposure", C.c_float)]


class FrameStats(C.Structure):
    _fields_ = [("extension_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("shaded_hits", C.c_uint64), ("frames", C.c_uint64),
                ("device_ms", C.c_float), ("kernel_launches", C.c_uint32)]


class KernelProfile(C.Structure):
    _fields_ = [("ms", C.c_float * 4), ("launches", C.c_uint32 * 4), ("closest_work", C.c_uint64 * 4), ("any_work", C.c_uint64 * 4),
                ("closest_sched", C.c_uint64 * 10), ("any_sched", C.c_uint64 * 10)]


assert C.sizeof(MaterialPod) == 92 and C.sizeof(Aabb) == 24

_lib = None


def lib():
    """Loads the CUDA library; fails loudly when it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NexusError(f"{LIB_PATH} is missing: build it with `make -C nexus_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.nx_last_error.restype = C.c_char_p
        L.nx_ctx_stream.restype = C.c_void_p
        if L.nx_abi_version() != 1:
            raise NexusError("libnexus_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(ctx_handle, rc, what):
    if rc < 0:
        msg = lib().nx_last_error(ctx_handle).decode() if ctx_handle else ""
        raise NexusError(f"{what} failed (status {rc}): {msg}")
    return rc
