#!/bin/bash
# every bench line of one round (run under gpurun, one GPU): scripts/bench_all.sh r01  ->  gpurun_out/bench_<tag>/*.json
TAG=${1:-r01}; OUT=gpurun_out/bench_$TAG; mkdir -p $OUT
for wl in instanced10m_4k cornell_1080p sky10m_4k build50m build10m build100k; do
  python bench.py --workload $wl --impl reference > $OUT/${wl}_reference.json 2> $OUT/${wl}_reference.err
  python bench.py --workload $wl > $OUT/${wl}_ours.json 2> $OUT/${wl}_ours.err
done
python - <<PY
import json, glob, os
for wl in ("instanced10m_4k", "cornell_1080p", "sky10m_4k", "build50m", "build10m", "build100k"):
    try:
        o = json.loads(open("$OUT/%s_ours.json" % wl).read().strip().splitlines()[-1]); r = json.loads(open("$OUT/%s_reference.json" % wl).read().strip().splitlines()[-1])
        print("%-16s ours %9.1f (e2e %9.1f) reference %9.1f %s  -> x%.2f (e2e x%.2f)  %.3f ms/step" % (wl, o["value"], o["e2e"]["value"], r["value"], o["unit"], o["value"] / r["value"], o["e2e"]["value"] / r["value"], o["ms_per_step"]))
    except Exception as e:
        print(wl, "failed:", e)
PY
