set -x
NX_FRAMES=4 timeout 600 python scripts/tune_pool.py instanced10m_4k lane:6,8 lane:6,4 lane:6,2 lane:6,1 lane:8,4 lane:10,4 lane:12,4 lane:4,4 lane:8,2 lane:10,2 lane:6,8 2>&1 | grep -v "^      any"
