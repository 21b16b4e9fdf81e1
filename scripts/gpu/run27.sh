set -x
timeout 900 python -m pytest tests/test_gpu_builder.py tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -25
