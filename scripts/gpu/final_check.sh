set -x
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/bench_final_ours.json; python scripts/jl.py value ms_per_step e2e.value roofline.frac roofline.bound gpu_launches wall_s < gpurun_out/bench_final_ours.json
timeout 400 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_final_ref.json; python scripts/jl.py value ms_per_step impl < gpurun_out/bench_final_ref.json
