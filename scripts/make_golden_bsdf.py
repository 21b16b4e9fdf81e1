"""Golden vectors of the reference's own BSDF code: D_PrincipledBSDF::Eval (Nexus/src/Cuda/BSDF/*.cuh, device-only in the reference)
compiled UNMODIFIED for the host with g++ (`make -C oracle refcpu`, shims for the device intrinsics in oracle/ref/ref_cpu_bsdf.cpp) and run
on the CPU on random (material, wi, wo) triples.   python scripts/make_golden_bsdf.py  ->  tests/golden/bsdf_ref.npz"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_lib as O
from golden_cases import bsdf_cases

assert O.have_refcpu(), "build oracle/_ref/libnexus_refcpu.so first: make -C oracle refcpu (needs /root/reference)"
R = C.CDLL(O.REFCPU_SO)
P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
mat, wi, wo = bsdf_cases()
n = len(mat)
f, pdf, ok = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
R.ref_bsdf_eval(P(mat), P(wi), P(wo), C.c_uint32(n), P(f), P(pdf), P(ok))
path = os.path.join(ROOT, "tests", "golden", "bsdf_ref.npz")
np.savez_compressed(path, bsdf=f, pdf=pdf, ok=ok)
print(path, os.path.getsize(path), "bytes;", n, "evaluations,", int(ok.sum()), "valid")
