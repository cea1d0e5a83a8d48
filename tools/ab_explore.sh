#!/bin/bash
# developer aid (GPU box): the same explore subset under the in-tree library and under a variant build, same box.
#   gpurun --timeout 600 -- 'bash tools/ab_explore.sh phased_v6 [ncu-variant-name]'
set -u
mkdir -p gpurun_out
sel=${1:-phased_v6}
for lib in default variant; do
  if [ $lib = variant ]; then export RENDU_B200_LIB=$PWD/rendu_b200/librendu_b200_variant.so; [ -f "$RENDU_B200_LIB" ] || continue; fi
  echo "== $lib"
  RB_EXPLORE_ONLY=$sel timeout 300 python bench.py --explore 2>&1 | grep "\[explore\]" | tee gpurun_out/ab_${lib}.log
done
unset RENDU_B200_LIB
if [ -n "${2:-}" ]; then
  RB_EXPLORE_ONLY=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rb_k_trace -c 4 -o gpurun_out/ab_ncu -f python bench.py --explore > gpurun_out/ab_ncu.log 2>&1
fi
