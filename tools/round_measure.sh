#!/bin/bash
# Round-end measurement pass on one B200: bench lines + ncu launch lists -> gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench_ref.err
python bench.py > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
python bench.py --spp 256 --steps 3 --warmup 3 > gpurun_out/m_pt_256.json 2> gpurun_out/m_pt_256.err
python bench.py --spp 256 --steps 3 --warmup 3 --trace-mode ordered --no-cpu > gpurun_out/m_pt_256_ordered.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_launches_pt.csv python bench.py --spp 16 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
for f in gpurun_out/m_bench_ref.json gpurun_out/m_bench.json gpurun_out/m_pt_256.json gpurun_out/m_pt_256_ordered.json; do cut -c1-400 $f; done
