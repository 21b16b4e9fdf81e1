set -x
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_scale.py 2>&1 | tail -6
for w in build10m build100k build50m; do echo "== main $w"; timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value ms_per_step roofline.stage_ms sah_optimal_collapse.total_ms sah_optimal_collapse.bvh8_ms morton64.stage_ms.sort_ms e2e.value e2e.unpipelined.value; done
echo "== NX_DP_WAVES=0 build10m"; NX_DP_WAVES=0 timeout 120 python bench.py --workload build10m --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py sah_optimal_collapse.total_ms sah_optimal_collapse.bvh8_ms
echo "== NX_COLLAPSE_CTA=0 build100k"; NX_COLLAPSE_CTA=0 timeout 120 python bench.py --workload build100k --steps 20 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py value ms_per_step roofline.stage_ms
echo "== cub"; for w in build10m build100k build50m; do NX_SORT=0 NEXUS_B200_LIB=$PWD/nexus_b200/variants/lib_cub.so timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-ncu 2>&1 | python scripts/jl.py roofline.stage_ms.sort_ms morton64.stage_ms.sort_ms; done
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
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -12
