set -x
timeout 900 python -m pytest tests/test_gpu_builder.py -m gpu -x -q 2>&1 | tail -4
for c in 400000 0; do
  echo "== cluster max $c"
  NX_COLLAPSE_CLUSTER_MAX=$c timeout 300 python bench.py --workload build100k --no-cpu-baseline --steps 16 --warmup 4 2>/dev/null | python scripts/jl.py value ms_per_step roofline.stage_ms sah_optimal_collapse.total_ms sah_optimal_collapse.bvh8_ms
done
python - <<PY
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, nexus_b200 as nx
from nexus_b200 import scenes
ctx = nx.Context(0)
for n in (60000, 100000, 200000, 400000, 800000, 1600000):
    prims = scenes.test_triangles(n)
    dev = ctx.upload(prims)
    row = []
    for cmax in (0, 10**9):
        os.environ['NX_COLLAPSE_CLUSTER_MAX'] = str(cmax)
        c2 = nx.Context(0)
        d2 = c2.upload(prims)
        ms = [nx.BenchmarkBuild(c2, d2, n, 1, True, 3, 10)['bvh8_ms'] for _ in range(2)]
        mo = [nx.BenchmarkBuild(c2, d2, n, 1, True, 3, 10, collapse=nx.COLLAPSE_SAH_OPTIMAL, maxLeafPrims=2)['bvh8_ms'] for _ in range(2)]
        row.append((min(ms), min(mo)))
        c2.free(d2); c2.close()
    print(n, 'grid: %.4f / opt %.4f   cluster: %.4f / opt %.4f' % (row[0][0], row[0][1], row[1][0], row[1][1]), flush=True)
PY
