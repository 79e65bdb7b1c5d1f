#!/bin/bash
# tuning sweep on the GPU box: tools/sweep2.sh "lib|bench flags" ...   (lib = default or a variant build under the package directory)
cd /root/repo
for cfg in "$@"; do
  IFS='|' read lib flags <<< "$cfg"
  if [ "$lib" = "default" ]; then unset ZKV_LIB; else export ZKV_LIB=/root/repo/stylus_zkvm_verifiers_b200/$lib; fi
  python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline $flags 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$cfg', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 2), {k: round(v, 2) for k, v in d['stage_ms'].items()}, 'wave', d['roofline']['launch']['proofs'], round(d['roofline']['launch']['ms'], 2), d['roofline']['final_exp_launch']['proofs'], round(d['roofline']['final_exp_launch']['ms'], 2))"
done
