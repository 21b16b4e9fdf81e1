set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_cpp_host.py tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | tail -12
./examples/render_multigpu 2 3840 2160 8 2>&1 | tail -3
