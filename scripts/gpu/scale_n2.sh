set -x
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded_build.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_inst_r02c_n2.json
python scripts/jl.py value ms_per_step scene_setup_s e2e.value reduce_ms reduce_check config5 < gpurun_out/bench_inst_r02c_n2.json
