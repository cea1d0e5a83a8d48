#!/bin/bash
# developer aid: end-to-end (host pointer) throughput of the raycaster bench for ring settings (GPU box)
run() { echo "== $*"; env "$@" python bench.py --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['e2e']['value'], d['value'])"; }
run A=1
run RB_PIPE_CHUNK=1048576
run RB_PIPE_CHUNK=262144 RB_PIPE_SLOTS=4
run RB_PIPE_CHUNK=1048576 RB_PIPE_SLOTS=4
run RB_PIPE_SLOTS=2
run RB_PIPE_CHUNK=2097152 RB_PIPE_SLOTS=2
