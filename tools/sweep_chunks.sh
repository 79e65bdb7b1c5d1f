#!/bin/bash
# bench.py at several batch sizes and stream-overlap chunk counts (A/B for the automatic chunking policy): sweep_chunks.sh N c1 c2 ...
n=$1; shift
for c in "$@"; do
  python bench.py --n $n --chunks $c --steps 4 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n', $n, 'chunks', $c, 'ms_per_step', round(d['ms_per_step'], 3), 'Mverifies/s', round(d['value'] / 1e6, 3), 'e2e_M', round(d['e2e']['value'] / 1e6, 3))"
done
