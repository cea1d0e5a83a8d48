#!/bin/bash
# developer aid: path-tracer throughput for lane / batch-size settings (run on the GPU box)
run() { echo "== $*"; env "$@" python bench.py --workload pathtrace --spp 64 --steps 3 --warmup 3 --no-cpu $EXTRA 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))"; }
EXTRA="" run RB_PT_LANES=2
EXTRA="" run RB_PT_LANES=2 RB_NODE_STEPS=1 RB_LEAF_VOTES=8 RB_FETCH_THRESHOLD=20
EXTRA="" run RB_PT_LANES=2 RB_NODE_STEPS=3
EXTRA="" run RB_PT_LANES=2 RB_NODE_STEPS=4 RB_FETCH_THRESHOLD=16
EXTRA="" run RB_PT_LANES=2 RB_NODE_STEPS=6 RB_FETCH_THRESHOLD=16 RB_LEAF_VOTES=4
EXTRA="" run RB_PT_LANES=2 RB_NODE_STEPS=6 RB_FETCH_THRESHOLD=8
EXTRA="--trace-mode ordered" run RB_PT_LANES=2
