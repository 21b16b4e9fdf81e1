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
"
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -8
