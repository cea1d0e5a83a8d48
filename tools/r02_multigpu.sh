#!/bin/bash
# Round-2 multi-GPU pass on one box with N GPUs: the driver's torchrun launch of bench.py (both arms' product side), the
# in-process multi-GPU tests, and the C++ drop-in at 1..N GPUs.   gpurun --gpus 8 --timeout 900 -- 'bash tools/r02_multigpu.sh'
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench N=$N rc=$?"
timeout -k 5 300 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -2
bash tools/dropin_scaling.sh 2>&1 | tail -9
