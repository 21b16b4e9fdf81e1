#!/bin/bash
# compute-sanitizer passes over the small GPU tests (run under gpurun, one GPU):  scripts/sanitize.sh [memcheck|racecheck|initcheck|synccheck]
# The persistent traversal kernels spin on device-side queue counters by design; racecheck therefore only covers shared memory
# (its scope anyway).  Output: gpurun_out/sanitize_<tool>.log ; exit code of the last run.
TOOL=${1:-memcheck}; OUT=gpurun_out; mkdir -p $OUT
export NX_SANITIZE=1
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_trace.py::test_edge_cases tests/test_gpu_trace.py::test_exact_ties_and_degenerate_triangles \
                   "tests/test_gpu_trace.py::test_every_loop_and_scene_kind_gives_the_same_bytes" \
                   tests/test_gpu_builder.py::test_invalid_inputs_fail_loudly tests/test_gpu_builder.py::test_refit_keeps_the_topology_and_reproduces_a_build_on_unchanged_boxes \
                   tests/test_gpu_render.py::test_present_is_a_pipelined_read_rgba8 tests/test_gpu_render.py::test_present_device_writes_the_display_image_into_caller_memory \
                   tests/test_gpu_render.py::test_tlas_refit_gives_the_hits_of_a_rebuilt_scene tests/test_gpu_render.py::test_dynamic_scene_updates_rebuild_the_tlas \
                   tests/test_gpu_render.py::test_pixel_query_returns_the_primary_hit_instance tests/test_gpu_host_api.py -x -q > $OUT/sanitize_$TOOL.log 2>&1
rc=$?
tail -n 15 $OUT/sanitize_$TOOL.log
exit $rc
