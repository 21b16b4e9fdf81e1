set -x
timeout 240 python -m pytest tests/test_gpu_builder.py -m gpu -x -q 2>&1 | tail -6
timeout 240 python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py tests/test_gpu_host_api.py tests/test_gpu_sharded_build.py -m gpu -x -q 2>&1 | tail -8
NX_TRACE_MODE=duo timeout 200 python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -8
NX_FRAMES=3 timeout 200 python scripts/tune_pool.py instanced10m_4k lane:6,8 duo:6,8 duo:10,10 duo:14,12 duo:4,6 lane:6,8 2>&1 | tail -24 | tee gpurun_out/tune_duo_2.log
for h in 0 2; do for so in 0 1; do echo "== NX_HPLOC=$h NX_SORT=$so build10m"; NX_HPLOC=$h NX_SORT=$so timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value ms_per_step roofline.stage_ms sah_optimal_collapse.total_ms sah_optimal_collapse.bvh8_ms morton64.stage_ms; done; done
for so in 0 1; do echo "== NX_SORT=$so build100k"; NX_SORT=$so timeout 120 python bench.py --workload build100k --steps 20 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value ms_per_step roofline.stage_ms sah_optimal_collapse.total_ms; done
for v in dp32 dp64; do echo "== $v build10m"; NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_$v.so timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value ms_per_step roofline.stage_ms sah_optimal_collapse.total_ms sah_optimal_collapse.bvh8_ms; done
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -8
