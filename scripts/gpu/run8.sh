set -x
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -25
for v in pb64 pb64r40 pb128r40 pb128r32; do echo "== $v"; for w in build10m build100k; do NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_$v.so timeout 120 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py roofline.stage_ms.bvh2_ms roofline.stage_ms.sort_ms value; done; done
echo "== main"; for w in build10m build100k; do timeout 120 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py roofline.stage_ms value ms_per_step; done
echo "== cub"; NX_SORT=0 NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_cub.so timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | tail -3 | cut -c1-600
mkdir -p gpurun_out/prof_r02b
BCMD="python bench.py --workload build10m --steps 1 --warmup 3 --no-cpu-baseline --no-ncu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/prof_r02b/launches_build10m.csv $BCMD > gpurun_out/prof_r02b/launches_build10m.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02b/onesweep_build10m $BCMD > gpurun_out/prof_r02b/full_onesweep.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hploc_seed_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02b/hploc_build10m $BCMD > gpurun_out/prof_r02b/full_hploc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dp_eval_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02b/dp_eval_build10m $BCMD > gpurun_out/prof_r02b/full_dp.log 2>&1
NX_SORT=0 NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_cub.so timeout 300 ncu --set full --clock-control none -k regex:Onesweep -s 4 -c 1 -f -o gpurun_out/prof_r02b/cub_onesweep_build10m $BCMD > gpurun_out/prof_r02b/full_cub.log 2>&1
ls -la gpurun_out/prof_r02b
