set -x
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -12
NX_FRAMES=4 timeout 200 python scripts/tune_pool.py instanced10m_4k lane:6,8 2>&1 | tail -4
NX_MERGE_INSTANCES=0 NX_FRAMES=4 timeout 200 python scripts/tune_pool.py instanced10m_4k lane:6,8 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -12
timeout 400 python bench.py --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_inst_r02b.json; python scripts/jl.py value ms_per_step scene_setup_s e2e.value roofline.kernel_ms_per_step roofline.frac roofline.lanes_per_inst roofline.per_ray two_level.value like_for_like.value cpu_baseline.value < gpurun_out/bench_inst_r02b.json
timeout 300 python bench.py --impl reference --steps 8 --warmup 3 2>&1 | tail -1 | python scripts/jl.py value ms_per_step
