#!/bin/bash
# compute-sanitizer passes over the small GPU tests (run under gpurun, one GPU):  scripts/sanitize.sh [memcheck|racecheck|initcheck|synccheck]
# The persistent traversal kernels spin on device-side queue counters by design; racecheck therefore only covers shared memory
# (its scope anyway).  Output: gpurun_out/sanitize_<tool>.log ; exit code of the last run.
TOOL=${1:-memcheck}; OUT=gpurun_out; mkdir -p $OUT
export NX_SANITIZE=1
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_trace.py::test_edge_cases tests/test_gpu_trace.py::test_exact_ties_and_degenerate_triangles \
                   tests/test_gpu_builder.py::test_invalid_inputs_fail_loudly tests/test_gpu_render.py::test_present_is_a_pipelined_read_rgba8 \
                   tests/test_gpu_render.py::test_pixel_query_returns_the_primary_hit_instance tests/test_gpu_host_api.py -x -q > $OUT/sanitize_$TOOL.log 2>&1
rc=$?
tail -n 15 $OUT/sanitize_$TOOL.log
exit $rc
