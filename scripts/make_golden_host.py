"""Golden vectors of the reference's own HOST code for the records the scene layer hands to the kernels: MeshInstance::ToDevice
(Nexus/src/Scene/MeshInstance.h:36-66) and Camera::ToDevice (Nexus/src/Scene/Camera.cpp:130-156), compiled unmodified with g++
(`make -C oracle refcpu`, harness oracle/ref/ref_cpu_host.cpp); runs in the build container, no GPU.
    python scripts/make_golden_host.py  ->  tests/golden/host_ref.npz"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_lib as O
from golden_cases import host_cases

assert O.have_refcpu(), "build oracle/_ref/libnexus_refcpu.so first: make -C oracle refcpu (needs /root/reference)"
R = C.CDLL(O.REFCPU_SO)
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
inst, cams = host_cases()
inst_out = np.zeros((len(inst["position"]), 160), np.uint8)
for k in range(len(inst_out)):
    R.ref_host_instance(P(inst["position"][k]), P(inst["rotation"][k]), P(inst["scale"][k]), P(inst["mesh_bounds"][k]),
                        C.c_uint32(int(inst["mesh_idx"][k])), C.c_uint32(int(inst["material_idx"][k])), P(inst_out[k]))
cam_out = np.zeros((len(cams["position"]), 88), np.uint8)
for k in range(len(cam_out)):
    R.ref_host_camera(P(cams["position"][k]), P(cams["forward"][k]), C.c_float(float(cams["hfov"][k])), C.c_float(float(cams["focus"][k])),
                      C.c_float(float(cams["defocus"][k])), C.c_uint32(int(cams["res"][k, 0])), C.c_uint32(int(cams["res"][k, 1])), P(cam_out[k]))
cam_out[:, 76:80] = 0      # padding before the 8-byte aligned resolution
path = os.path.join(ROOT, "tests", "golden", "host_ref.npz")
np.savez_compressed(path, instance_records=inst_out, camera_records=cam_out)
print(path, os.path.getsize(path), "bytes;", len(inst_out), "instances,", len(cam_out), "cameras")
