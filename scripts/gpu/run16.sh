set -x
for m in "NX_MERGE_MIN_PRIMS=0" "NX_MERGE_MIN_PRIMS=256" "NX_MERGE_INSTANCES=0"; do echo "== $m"; env NX_DEBUG=1 $m timeout 250 python scripts/diag_merged.py 2>&1 | tail -9; done
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -6
