set -x
OUT=gpurun_out/prof_r02f; mkdir -p $OUT
BCMD="python bench.py --workload build10m --steps 1 --warmup 3 --no-cpu-baseline --no-ncu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_build10m.csv $BCMD > $OUT/launches_build10m.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hploc_seed_kernel -s 3 -c 1 -f -o $OUT/hploc_build10m $BCMD > $OUT/full_hploc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:collapse_kernel -s 3 -c 1 -f -o $OUT/collapse_build10m $BCMD > $OUT/full_collapse.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 4 -c 1 -f -o $OUT/onesweep_build10m $BCMD > $OUT/full_onesweep.log 2>&1
ls -la $OUT
