"""Every `File.ext:line[-line]` citation of the reference in the headers, the host layers, the oracle and the design documents must
resolve: a file of that name exists under /root/reference (the longest path suffix given must match) and has at least that many lines.
Guards the citations the parity review relies on against rot.  Skipped where the reference is not mounted (the GPU box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SOURCES = ["include/nexus_b200.h", "include/nexus_b200.hpp", "include/nexus_b200_import.hpp", "INTEGRATION.md", "DESIGN.md", "nexus_b200/__init__.py",
           "nexus_b200/gltf.py", "nexus_b200/obj.py", "nexus_b200/hdr.py", "nexus_b200/multigpu.py", "nexus_b200/csrc/traverse.cuh", "nexus_b200/csrc/bsdf.cuh",
           "nexus_b200/csrc/render.cu", "nexus_b200/csrc/shade.cu", "nexus_b200/csrc/scene.cu", "nexus_b200/csrc/bvh_builder.cu", "nexus_b200/csrc/wave.cuh",
           "oracle/oracle_bvh.cpp", "oracle/oracle_trace.cpp", "oracle/oracle_sah.cpp", "oracle/oracle_common.h", "oracle/ref/ref_cpu_collapse.cpp",
           "oracle/ref/ref_cpu_host.cpp", "oracle/ref/ref_cpu_bsdf.cpp"]
CITE = re.compile(r"((?:[A-Za-z0-9_.]+/)*[A-Za-z0-9_]+\.(?:cuh|cu|cpp|h|hpp|md|txt)):(\d+)(?:-(\d+))?")
OURS = {"nexus_b200.h", "nexus_b200.hpp", "traverse.cuh", "bsdf.cuh", "wave.cuh", "scene.cuh", "nx_common.cuh", "render.cu", "shade.cu", "scene.cu", "bvh_builder.cu",
        "context.cu", "oracle_sah.cpp", "oracle_bvh.cpp", "oracle_trace.cpp", "oracle_common.h", "SURVEY.md", "DESIGN.md", "INTEGRATION.md", "BASELINE.md"}
ALIAS = {"N": "Nexus/src", "B": "Nexus/vendor/NexusBVH/NexusBVH", "T": "Nexus/vendor/NexusBVH/Test", "src": "src", "vendor": "vendor"}


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the build container")
def test_reference_citations_resolve():
    files = {}
    for d, _, names in os.walk(REF):
        if "/vendor/assimp" in d or "/vendor/glm" in d or "/.git" in d:
            continue
        for n in names:
            files.setdefault(n, []).append(os.path.join(d, n))
    lines = {}
    bad, checked = [], 0
    for src in SOURCES:
        text = open(os.path.join(ROOT, src), errors="replace").read()
        for m in CITE.finditer(text):
            path, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            base = os.path.basename(path)
            if base in OURS or base not in files:
                if base not in OURS and base not in files:
                    bad.append(f"{src}: {m.group(0)} - no file named {base} in the reference")
                continue
            parts = path.split("/")
            if parts[0] in ALIAS:
                parts = ALIAS[parts[0]].split("/") + parts[1:]
            suffix = "/".join(parts)
            cands = [f for f in files[base] if f.endswith("/" + suffix)] or [f for f in files[base] if f.endswith("/" + "/".join(parts[-2:]))] or files[base]
            ok = False
            for f in cands:
                if f not in lines:
                    lines[f] = sum(1 for _ in open(f, errors="replace"))
                ok = ok or (1 <= lo <= hi <= lines[f])
            checked += 1
            if not ok:
                bad.append(f"{src}: {m.group(0)} - {os.path.relpath(cands[0], REF)} has {lines[cands[0]]} lines")
    assert checked > 200, checked
    assert not bad, "\n".join(bad)
