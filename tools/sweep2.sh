cd /root/repo
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$*', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 2), d['gpu_launches'])"; }
run --chunks 2 --segments 8
run --chunks 2 --segments 16
run --chunks 3 --segments 8
run --chunks 4 --segments 8
run --chunks 4 --segments 16
run --chunks 2 --segments 4
run --n 131072
run --n 262144
