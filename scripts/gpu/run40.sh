set -x
OUT=gpurun_out/prof_r02f; mkdir -p $OUT
BCMD="python bench.py --workload build10m --steps 1 --warmup 3 --no-cpu-baseline --no-ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dp_wave_kernel -s 40 -c 24 -f -o $OUT/dp_wave_build10m $BCMD > $OUT/full_dpwave.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dp_ -c 200 --csv --log-file $OUT/launches_dp_build10m.csv $BCMD > $OUT/launches_dp.log 2>&1
ls -la $OUT | tail -5
