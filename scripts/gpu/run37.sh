for v in "" pb32 pb64 pb256 ""; do
  echo "== variant '$v'"
  if [ -n "$v" ]; then export NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_$v.so; else unset NEXUS_B200_LIB; fi
  for w in build10m build100k; do
    timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 12 --warmup 4 2>/dev/null | python scripts/jl.py ms_per_step roofline.stage_ms.bvh2_ms roofline.stage_ms.total_ms
  done
done
