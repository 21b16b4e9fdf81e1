set -x
NX_HPLOC=2 timeout 240 python -m pytest tests/test_gpu_builder.py -m gpu -x -q 2>&1 | tail -6
timeout 240 python -m pytest tests/test_gpu_builder.py -m gpu -x -q 2>&1 | tail -3
for h in 0 2; do echo "== NX_HPLOC=$h build10m"; NX_HPLOC=$h timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value ms_per_step roofline.stage_ms sah_optimal_collapse.total_ms morton64.stage_ms; done
for v in sort256 sort384; do echo "== $v build10m"; NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_$v.so timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py roofline.stage_ms morton64.stage_ms; done
echo "== build100k"; NX_HPLOC=2 timeout 120 python bench.py --workload build100k --steps 20 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value ms_per_step roofline.stage_ms
echo "== build50m"; NX_HPLOC=2 timeout 200 python bench.py --workload build50m --steps 3 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value ms_per_step roofline.stage_ms sah_optimal_collapse.total_ms e2e.value
timeout 400 python bench.py --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_inst_r02a.json; python scripts/jl.py value ms_per_step scene_setup_s e2e.value roofline.kernel_ms_per_step roofline.frac roofline.traffic roofline.lanes_per_inst roofline.issue_slot_util like_for_like.value cpu_baseline.value < gpurun_out/bench_inst_r02a.json
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -8
