"""Prints selected keys of the last JSON line on stdin: python scripts/jl.py key1 key2.sub ..."""
import json, sys
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not line:
    print("no json line"); sys.exit(0)
d = json.loads(line[-1])
out = {}
for k in sys.argv[1:]:
    v = d
    for p in k.split("."):
        v = v.get(p) if isinstance(v, dict) else None
    out[k] = v
print(json.dumps(out))
