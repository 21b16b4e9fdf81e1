set -x
timeout 600 python -m pytest tests/test_gpu_trace.py -m gpu -x -q 2>&1 | tail -12
NX_FRAMES=4 timeout 300 python scripts/tune_pool.py instanced10m_4k lane lane 2>&1 | grep -v "^      any"
NX_MERGE_INSTANCES=0 NX_FRAMES=4 timeout 300 python scripts/tune_pool.py instanced10m_4k lane 2>&1 | grep -v "^      any"
