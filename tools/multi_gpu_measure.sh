#!/bin/bash
# usage: tools/multi_gpu_measure.sh N   — raycaster bench + 1080p/1024spp frame (tile and sample split) on N GPUs of one box
N=$1
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --no-cpu > gpurun_out/mg_bench_${N}gpu.json 2> gpurun_out/mg_bench_${N}gpu.err
run --spp 1024 --split tile --steps 3 --warmup 3 --no-cpu > gpurun_out/mg_pt_${N}gpu_tile.json 2> gpurun_out/mg_pt_${N}gpu_tile.err
run --spp 1024 --split sample --steps 3 --warmup 3 --no-cpu > gpurun_out/mg_pt_${N}gpu_sample.json 2> gpurun_out/mg_pt_${N}gpu_sample.err
for f in gpurun_out/mg_bench_${N}gpu.json gpurun_out/mg_pt_${N}gpu_tile.json gpurun_out/mg_pt_${N}gpu_sample.json; do grep '^{' $f | cut -c1-160; done
