set -x
timeout 900 python -m pytest tests/test_cpp_host.py tests/test_gpu_host_api.py -m gpu -x -q 2>&1 | tail -15
