set -x
TAG=r02c; WL=instanced10m_4k
OUT=gpurun_out/prof_$TAG; mkdir -p $OUT
NX_FRAMES=4 timeout 300 python scripts/tune_pool.py instanced10m_4k lane lane:6,8 lane:8,4 lane:8,6 lane:7,4 lane:8,3 lane 2>&1 | grep -v "^      any" > $OUT/tune_direct.log
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-like-for-like --no-ncu"
K='regex:trace_closest_kernel|trace_any_kernel|shade_kernel|generate_kernel|frame_totals_kernel|resolve_rgba8_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 140 --csv --log-file $OUT/launches_$WL.csv $CMD > $OUT/launches_$WL.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_closest_kernel -s 32 -c 3 -f -o $OUT/trace_closest_$WL $CMD > $OUT/full_closest_$WL.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_any_kernel -s 33 -c 1 -f -o $OUT/trace_any_$WL $CMD > $OUT/full_any_$WL.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 33 -c 2 -f -o $OUT/shade_$WL $CMD > $OUT/full_shade_$WL.log 2>&1
ls -la $OUT
cat $OUT/tune_direct.log
