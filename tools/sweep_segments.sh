#!/bin/bash
# bench.py at several Miller-segment / final-exponentiation-stage settings (A/B of the chunk interleaving grain)
for cfg in "$@"; do
  seg=${cfg%%:*}; fe=${cfg##*:}
  python bench.py --segments $seg --fe-stages $fe --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('segments', $seg, 'fe_stages', $fe, 'ms_per_step', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'])"
done
