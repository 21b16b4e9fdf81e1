set -x
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
timeout 400 python bench.py --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_inst_r02c.json; python scripts/jl.py value ms_per_step scene_setup_s e2e.value roofline.kernel_ms_per_step roofline.frac roofline.lanes_per_inst roofline.per_ray two_level.value like_for_like.value cpu_baseline.value < gpurun_out/bench_inst_r02c.json
timeout 300 python bench.py --impl reference --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_inst_r02c_ref.json; python scripts/jl.py value ms_per_step < gpurun_out/bench_inst_r02c_ref.json
