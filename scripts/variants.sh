#!/bin/bash
# runs scripts/tune_trace.py for every built kernel variant (GPU), REPS times each, interleaved
WL=${1:-instanced10m_4k}; REPS=${2:-2}
for r in $(seq $REPS); do
echo "== default"; python scripts/tune_trace.py $WL 6,6 2>&1 | tail -1
for v in nexus_b200/variants/lib_*.so; do echo "== $v"; NEXUS_B200_LIB=$PWD/$v python scripts/tune_trace.py $WL 6,6 2>&1 | tail -1; done
done
