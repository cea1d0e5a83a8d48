#!/bin/bash
# The eight-wide kernels on one B200 (first run: round 2): parity of the eight-wide kernels, their timing against the default
# scheduler on C2, and an ncu capture of the incoherent closest-hit launch of the best one. Outputs land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/wide_measure.sh'
set -u
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_wide.py tests/test_raycaster_gpu.py -q -m gpu > gpurun_out/wide_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/wide_tests.log
RB_EXPLORE_ONLY=phased_v6_t12_s6,wide timeout 300 python bench.py --explore 2>&1 | grep "\[explore\]" | tee gpurun_out/wide_explore.log
cp gpurun_out/explore.json gpurun_out/wide_explore.json 2>/dev/null
# Launch list and full capture of the wide kernel alone (first variant = the one profiled): never a bench number.
RB_EXPLORE_ONLY=wide_t12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rb_k_trace_wide -c 4 -o gpurun_out/wide_t12 -f python bench.py --explore > gpurun_out/wide_ncu.log 2>&1
RB_EXPLORE_ONLY=wide_phased_v6_t12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rb_k_trace_wide -c 4 -o gpurun_out/wide_phased_v6_t12 -f python bench.py --explore >> gpurun_out/wide_ncu.log 2>&1
# the path tracer's closest-hit wavefront on the wide tree against exact mode (16 spp frames, same box)
for m in exact wide; do timeout -k 5 200 python bench.py --spp 16 --steps 3 --warmup 2 --no-cpu --trace-mode $m 2>/dev/null | tail -1 > gpurun_out/wide_pt_$m.json; done
ls -la gpurun_out | tail -8
