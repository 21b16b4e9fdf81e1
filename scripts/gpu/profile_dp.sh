OUT=gpurun_out/prof_r02f; mkdir -p $OUT
BCMD="python bench.py --workload build10m --steps 1 --warmup 3 --no-cpu-baseline --no-ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dp_wave_kernel -s 32 -c 3 -f -o $OUT/dp_wave_build10m $BCMD > $OUT/full_dpwave.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dp_height_scatter_kernel -s 1 -c 1 -f -o $OUT/dp_scatter_build10m $BCMD > $OUT/full_dpscatter.log 2>&1
ls -la $OUT | tail -3
