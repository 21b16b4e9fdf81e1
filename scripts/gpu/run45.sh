for g in 64 32 128; do
  echo "== L2 fetch $g"
  for w in build10m build50m; do
    NX_L2_FETCH=$g timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 12 --warmup 4 2>/tmp/err.log | python scripts/jl.py ms_per_step roofline.stage_ms sah_optimal_collapse.total_ms
    grep "nx\] L2" /tmp/err.log | head -1
  done
  NX_L2_FETCH=$g NX_FRAMES=6 timeout 300 python scripts/tune_pool.py instanced10m_4k lane 2>&1 | grep -v "^      "
done
