#!/bin/bash
# compute-sanitizer passes over small product runs (GPU box). Output: gpurun_out/sanitize_*.log
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 compute-sanitizer --tool "$TOOL" --error-exitcode 9 --print-limit 20 "$@" > gpurun_out/sanitize_${TOOL}_${name}.log 2>&1; echo "$TOOL $name rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${TOOL}_${name}.log) $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${TOOL}_${name}.log | tail -1)"; }
if [ "${SEL:-}" = s3 ]; then
for TOOL in memcheck racecheck; do
run s3_pt python -m pytest tests/test_pathtracer_gpu.py -q -x -k "cp_async or exact_division or tiny_frames or path_for_path"
run s3_lbvh python -m pytest tests/test_lbvh_gpu.py -q -x -k "multi_mesh or degenerate or refit"
done
exit 0
fi
for TOOL in memcheck racecheck; do
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run golden python -m pytest tests/test_raycaster_gpu.py -q -x -k "golden or ray_levels or many_roots or single_ray"
run lbvh python -m pytest tests/test_lbvh_gpu.py -q -x -k "multi_mesh or refit"
run pt python -m pytest tests/test_pathtracer_gpu.py -q -x -k "path_for_path or output_stage or seed_determinism"
run r02 python -m pytest tests/test_pathtracer_gpu.py tests/test_raycaster_gpu.py tests/test_multigpu_gpu.py -q -x -k "live_session or recommitted or profile_frame or multi_gpu_render or primary_rays or per_triangle or empty_mesh or after_an_upload or garbage_records"
done
# session 3 of round 2: the cp.async-fed wavefront kernels (shared-memory staging: racecheck), the division KAT, the LBVH
# builder with collapsed leaves, the shade kernel after its diet.   SEL=s3 bash tools/sanitize.sh
