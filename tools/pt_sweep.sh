#!/bin/bash
# developer aid: path-tracer throughput for scheduler / lane / batch-size settings (GPU box), one short headline run each
run() { echo "== $EXTRA $*"; env "$@" python bench.py --spp 64 --steps 3 --warmup 3 --no-cpu --no-raycast $EXTRA 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['spp_per_s'],1), round(d['ms_per_step'],2))"; }
EXTRA="" run A=1
for t in 8 12 20 24; do EXTRA="" run RB_FETCH_THRESHOLD=$t; done
for n in 2 6 8; do EXTRA="" run RB_NODE_STEPS=$n; done
for v in 4 8 12; do EXTRA="" run RB_LEAF_VOTES=$v; done
EXTRA="" run RB_PT_LANES=3
EXTRA="--max-paths 33554432" run RB_PT_LANES=2
EXTRA="--max-paths 8388608" run RB_PT_LANES=2
EXTRA="" run A=2
