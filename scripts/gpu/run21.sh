set -x
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -15
