#!/bin/bash
# developer aid: A/B two builds of the library on the path-tracing workload (GPU box)
run() { echo "== $*"; env "$@" python bench.py --spp 64 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'])"; }
for i in 1 2; do
run A=1
run RENDU_B200_LIB=$PWD/rendu_b200/librendu_b200_variant.so
done
