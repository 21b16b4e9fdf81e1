set -x
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_host_api.py tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -8
for p in 2 1 2 1; do
NX_PIPES=$p NX_FRAMES=8 timeout 300 python scripts/tune_pool.py instanced10m_4k lane 2>&1 | grep -v "^      " | sed "s/^/pipes=$p /"
done
for p in 2 1; do
NX_PIPES=$p NX_FRAMES=16 timeout 300 python scripts/tune_pool.py cornell_1080p lane 2>&1 | grep -v "^      " | sed "s/^/pipes=$p /"
NX_PIPES=$p NX_FRAMES=8 timeout 300 python scripts/tune_pool.py sky10m_4k lane 2>&1 | grep -v "^      " | sed "s/^/pipes=$p /"
done
