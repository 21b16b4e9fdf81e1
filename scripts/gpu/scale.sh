N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 16 --warmup 4 2>&1 | tail -1 > gpurun_out/bench_inst_r02d_n$N.json
python scripts/jl.py value ms_per_step scene_setup_s e2e.value reduce_ms reduce_check.rel_err config5.value config5.spp_per_s config5.ms_per_frame config5.reduce_ms config5.reduce_ms_last_arrival config5.reduce_alone_ms config5.reduce_busbw_GBs < gpurun_out/bench_inst_r02d_n$N.json
