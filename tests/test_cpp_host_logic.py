"""Host logic of the C++ layer (include/nexus_b200.hpp) without a GPU: examples/host_logic_check.cpp is linked against a recording stub of
libnexus_b200.so generated from the header (every declared entry point logs its name and returns success), and the sequence of ABI calls
each step of the reference's edit pattern makes is checked - the C++ twin of tests/test_host_logic.py."""
import os
import subprocess

from test_abi import ROOT, declared_functions

SPECIAL = {
    "nx_abi_version": "int nx_abi_version(void) { return 1; }",
    "nx_last_error": "const char* nx_last_error(void* c) { (void)c; return \"stub\"; }",
    "nx_ctx_create": "int nx_ctx_create(int d, void** out) { (void)d; LOG(\"nx_ctx_create\"); *out = (void*)8; return 0; }",
    "nx_ctx_stream": "void* nx_ctx_stream(void* c) { (void)c; return 0; }",
    "nx_scene_create": "int nx_scene_create(void* c, unsigned w, unsigned h, void** out) { (void)c; (void)w; (void)h; LOG(\"nx_scene_create\"); *out = (void*)16; return 0; }",
    "nx_renderer_create": "int nx_renderer_create(void* c, unsigned w, unsigned h, void** out) { (void)c; (void)w; (void)h; LOG(\"nx_renderer_create\"); *out = (void*)24; return 0; }",
    "nx_ctx_destroy": "void nx_ctx_destroy(void* c) { (void)c; }", "nx_scene_destroy": "void nx_scene_destroy(void* c) { (void)c; }",
    "nx_renderer_destroy": "void nx_renderer_destroy(void* c) { (void)c; }",
}
COUNTED = {"nx_scene_add_material": "mat", "nx_scene_add_mesh": "mesh", "nx_scene_add_mesh_prebuilt": "mesh", "nx_scene_add_instance": "inst",
           "nx_scene_add_instance_matrix": "inst", "nx_scene_add_light": "light", "nx_scene_add_texture": "tex"}


def _stub_source():
    src = ["#include <stdio.h>", "#include <stdlib.h>",
           "static void LOG(const char* n) { const char* p = getenv(\"NX_STUB_LOG\"); if (!p) return; FILE* f = fopen(p, \"a\"); if (f) { fprintf(f, \"%s\\n\", n); fclose(f); } }",
           "void nx_stub_mark(const char* label) { const char* p = getenv(\"NX_STUB_LOG\"); if (!p) return; FILE* f = fopen(p, \"a\"); if (f) { fprintf(f, \"# %s\\n\", label); fclose(f); } }",
           "static int mat, mesh, inst, light, tex;"]
    for name in declared_functions():
        if name in SPECIAL:
            src.append(SPECIAL[name])
        elif name in COUNTED:
            src.append(f"int {name}() {{ LOG(\"{name}\"); return {COUNTED[name]}++; }}")
        else:
            src.append(f"int {name}() {{ LOG(\"{name}\"); return 0; }}")      # unspecified parameters (C): callable with any arguments
    return "\n".join(src) + "\n"


def test_cpp_update_pushes_exactly_what_was_invalidated(tmp_path):
    (tmp_path / "stub.c").write_text(_stub_source())
    subprocess.check_call(["gcc", "-std=gnu11", "-w", "-fPIC", "-shared", str(tmp_path / "stub.c"), "-o", str(tmp_path / "libnexus_b200.so")])
    exe = str(tmp_path / "host_logic_check")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "host_logic_check.cpp"),
                           "-L" + str(tmp_path), "-lnexus_b200", "-Wl,-rpath," + str(tmp_path), "-o", exe])
    log = tmp_path / "calls.log"
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, NX_STUB_LOG=str(log)))
    assert r.returncode == 0 and "host logic ok" in r.stdout, (r.returncode, r.stderr)
    steps, cur = {}, "setup"
    for line in log.read_text().splitlines():
        if line.startswith("# "):
            cur = line[2:]; steps[cur] = []
        else:
            steps.setdefault(cur, []).append(line)
    noise = {"nx_scene_set_render_settings"}                          # the settings struct is copied on every Update (a host-side struct copy)
    strip = lambda calls: [c for c in calls if c not in noise]        # noqa: E731
    assert strip(steps["clean update"]) == ["nx_scene_update"]                     # nothing dirty
    assert steps["edits"] == []                                                    # edits alone make no ABI call
    assert strip(steps["update after edits"]) == ["nx_scene_set_camera", "nx_scene_set_material", "nx_scene_set_instance_transform",
                                                  "nx_scene_set_instance_material", "nx_scene_set_light", "nx_scene_update"]
    assert strip(steps["second update"]) == ["nx_scene_update"]                    # nothing is pushed twice
    assert strip(steps["remove light"]) == ["nx_scene_remove_light", "nx_scene_set_light", "nx_scene_update"]
    assert steps["set transform"] == ["nx_scene_set_instance_transform"]           # applied at once
