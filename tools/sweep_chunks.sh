#!/bin/bash
# bench.py at several stream-overlap chunk counts (A/B for the automatic half-wave policy)
for c in "$@"; do
  python bench.py --chunks $c --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('chunks', $c, 'ms_per_step', round(d['ms_per_step'], 3), 'e2e_M', round(d['e2e']['value'] / 1e6, 3))"
done
