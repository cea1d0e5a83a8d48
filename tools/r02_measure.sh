#!/bin/bash
# Round-2 measurement pass on one B200: GPU tests, the default bench line (both arms), launch lists and the ncu captures
# the roofline blocks cite. Outputs land in gpurun_out/.   gpurun --timeout 1500 -- 'bash tools/r02_measure.sh [quick]'
set -u
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest.log; tail -3 gpurun_out/r02_pytest.log
timeout -k 5 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench.err
timeout -k 5 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
[ "${1:-}" = quick ] && exit 0
# launch list of a short headline run (never a bench number)
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --spp 16 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
# full captures: the C2 launches, then the three big path-tracer kernels of a 16-spp frame
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:rb_k_trace_phased -s 13 -c 4 -o gpurun_out/r02_ncu_raycast -f python bench.py --workload raycast --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_ncu.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"rb_k_pt_(trace|shade|shadow)" -s 72 -c 9 -o gpurun_out/r02_ncu_pt -f python bench.py --spp 16 --steps 1 --warmup 3 --no-cpu --no-raycast >> gpurun_out/r02_ncu.log 2>&1
ls -la gpurun_out | tail -6
