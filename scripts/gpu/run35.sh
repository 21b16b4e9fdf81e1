set -x
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
