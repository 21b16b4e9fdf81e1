set -x
NX_FRAMES=4 timeout 300 python scripts/tune_pool.py instanced10m_4k lane lane 2>&1 | grep -v "^      "
NX_FRAMES=8 timeout 300 python scripts/tune_pool.py cornell_1080p lane lane 2>&1 | grep -v "^      "
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -5
