#!/bin/bash
# developer aid: path-tracer throughput for lane / batch-size / scheduler settings (GPU box)
run() { echo "== $EXTRA $*"; env "$@" python bench.py --spp 64 --steps 3 --warmup 3 --no-cpu $EXTRA 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))"; }
EXTRA="" run RB_PT_LANES=2
EXTRA="--max-paths 33554432" run RB_PT_LANES=2
EXTRA="--max-paths 8388608" run RB_PT_LANES=2
EXTRA="--max-paths 33554432" run RB_PT_LANES=3
EXTRA="" run RB_PT_LANES=2
EXTRA="--max-paths 33554432" run RB_PT_LANES=2
