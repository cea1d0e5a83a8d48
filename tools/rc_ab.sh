#!/bin/bash
# developer aid: A/B two builds of the library on the raycaster bench (GPU box)
run() { echo "== $*"; env "$@" python bench.py --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['per_launch'].items()})"; }
for i in 1 2; do
run A=1
run RENDU_B200_LIB=$PWD/rendu_b200/librendu_b200_variant.so
done
