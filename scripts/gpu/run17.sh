set -x
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k config2 2>&1 | tail -30
NX_FRAMES=4 timeout 200 python scripts/tune_pool.py instanced10m_4k lane:6,8 2>&1 | tail -4
NX_MERGE_INSTANCES=0 NX_FRAMES=4 timeout 200 python scripts/tune_pool.py instanced10m_4k lane:6,8 2>&1 | tail -4
