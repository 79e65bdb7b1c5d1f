#!/bin/bash
# tuning sweep on the GPU box: variant builds (ZKV_LIB) x overlap chunks x batch size; prints value / e2e / stage_ms per configuration
# usage: tools/sweep.sh lib:chunks[:n] ...
cd /root/repo
for cfg in "$@"; do
  IFS=: read lib chunks n <<< "$cfg"
  n=${n:-65536}
  if [ "$lib" = "default" ]; then unset ZKV_LIB; else export ZKV_LIB=/root/repo/$lib; fi
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --chunks $chunks --n $n 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); s = d['stage_ms']
        print('$cfg', round(d['value']), round(d['e2e']['value']), {k: round(v, 2) for k, v in s.items()}, round(d['ms_per_step'], 2), 'miller/s=%.0f fe/s=%.0f' % ($n / s['miller'] * 1e3, $n / s['final_exp'] * 1e3))"
done
