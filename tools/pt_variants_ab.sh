#!/bin/bash
# developer aid (GPU box): time every rendu_b200/librendu_b200_variant_*.so against the default build on the path-tracing
# frame, interleaved, and print spp/s, ms per frame and the image hash (variants must not change the image).
#   tools/pt_variants_ab.sh [rounds] [spp]
rounds=${1:-2}; spp=${2:-64}
run() { name=$1; shift; env "$@" python bench.py --spp $spp --steps 3 --warmup 3 --no-cpu --no-raycast 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$name', round(d['spp_per_s'],1), round(d['ms_per_step'],2), d['image']['sha256'][:16], d['clocks']['sm_mhz'], ' '.join(k+'=%.2f' % v['ms_per_profiled_frame'] for k,v in d['roofline']['kernels'].items()))"; }
for i in $(seq $rounds); do
run default A=1
for so in rendu_b200/librendu_b200_variant_*.so; do
  n=${so##*variant_}; n=${n%.so}
  run $n RENDU_B200_LIB=$PWD/$so
done
done
