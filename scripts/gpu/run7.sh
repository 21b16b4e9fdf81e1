set -x
NX_PROFILE_SETUP=1 timeout 200 python -c "
import time, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench, nexus_b200 as nx
from nexus_b200 import scenes
ctx = nx.Context(0)
desc = bench.make_desc('instanced10m_4k')
for rep in range(2):
    t = time.time(); scene = scenes.build(ctx, desc, (3840, 2160)); ctx.synchronize(); print('scene_setup_s', round(time.time() - t, 3), flush=True)
    scene.close()
" 2>&1 | grep -E "scene_setup|nx setup"
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -6
for v in i16m5 i20m3 i24m2 i32t128; do echo "== $v build10m"; NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_$v.so timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py roofline.stage_ms.sort_ms morton64.stage_ms.sort_ms; done
echo "== main build10m"; timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value roofline.stage_ms morton64.stage_ms.sort_ms e2e
echo "== cub build10m"; NX_SORT=0 NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_cub.so timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py roofline.stage_ms.sort_ms morton64.stage_ms.sort_ms
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -8
