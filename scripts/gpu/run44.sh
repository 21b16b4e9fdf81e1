set -x
timeout 900 python -m pytest tests/test_gpu_builder.py -m gpu -x -q 2>&1 | tail -4
for w in build10m build50m build100k; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 12 --warmup 4 2>/dev/null | python scripts/jl.py ms_per_step roofline.stage_ms.bvh8_ms roofline.stage_ms.total_ms sah_optimal_collapse.total_ms sah_optimal_collapse.bvh8_ms
done
