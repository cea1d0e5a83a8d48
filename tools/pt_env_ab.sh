#!/bin/bash
# developer aid (GPU box): A/B of environment switches of the default library on the path-tracing frame
#   tools/pt_env_ab.sh "A=1" "RB_PT_FEED=1" ...   (each argument: one environment setting, run interleaved, two rounds)
spp=${SPP:-64}
run() { env "$@" python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu --no-raycast 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', round(d['spp_per_s'],1), round(d['ms_per_step'],2), d['image']['sha256'][:16], ' '.join(k+'=%.2f' % v['ms_per_profiled_frame'] for k,v in d['roofline']['kernels'].items()))"; }
for i in 1 2; do for e in "$@"; do run $e; done; done
