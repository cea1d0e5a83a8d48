#!/bin/bash
# developer aid: end-to-end (host pointer) throughput of the raycast block for ring settings (GPU box)
run() { echo "== $*"; env "$@" python bench.py --workload raycast --no-cpu --steps 6 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['e2e']['value']), round(d['e2e']['with_device_generated_primary_rays']['value']), round(d['value']))"; }
run A=1
run RB_PIPE_CHUNK=1048576
run RB_PIPE_CHUNK=2097152
run RB_PIPE_CHUNK=262144 RB_PIPE_SLOTS=4
run RB_PIPE_CHUNK=524288 RB_PIPE_SLOTS=4
run RB_PIPE_CHUNK=1048576 RB_PIPE_SLOTS=4
run RB_PIPE_CHUNK=262144 RB_PIPE_SLOTS=6
run RB_PIPE_CHUNK=524288 RB_PIPE_SLOTS=2
run RB_BENCH_E2E_THREADS=1
