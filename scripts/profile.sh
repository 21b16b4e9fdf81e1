#!/bin/bash
# ncu evidence for one round (run under gpurun, one GPU):  scripts/profile.sh r01 [workload]
# 1. launch list of the render kernels with device times (shares, not absolutes: cold cache, serialised)
# 2. --set full capture of the traversal kernels (primary-ray launch and a deep-bounce launch), the shadow kernel and shade
# 3. launch list + --set full capture of the builder kernels (H-PLOC, collapse) on the 10M-triangle benchmark mesh
TAG=${1:-r01}; WL=${2:-instanced10m_4k}
OUT=gpurun_out/prof_$TAG; mkdir -p $OUT
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-like-for-like"
K='regex:trace_closest_kernel|trace_any_kernel|shade_kernel|generate_kernel|frame_totals_kernel|resolve_rgba8_kernel'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 140 --csv --log-file $OUT/launches_$WL.csv $CMD > $OUT/launches_$WL.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_closest_kernel -s 32 -c 3 -f -o $OUT/trace_closest_$WL $CMD > $OUT/full_closest_$WL.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_any_kernel -s 33 -c 1 -f -o $OUT/trace_any_$WL $CMD > $OUT/full_any_$WL.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 33 -c 1 -f -o $OUT/shade_$WL $CMD > $OUT/full_shade_$WL.log 2>&1
BCMD="python bench.py --workload build10m --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_build10m.csv $BCMD > $OUT/launches_build10m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hploc_kernel -s 3 -c 1 -f -o $OUT/hploc_build10m $BCMD > $OUT/full_hploc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:collapse_kernel -s 3 -c 1 -f -o $OUT/collapse_build10m $BCMD > $OUT/full_collapse.log 2>&1
ls -la $OUT
