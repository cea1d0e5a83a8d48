#!/bin/bash
# The C++ drop-in (reference engine + GPU classes) on 1 / 2 / 4 / 8 GPUs of one box, both splits: rendering time of the
# second frame (the first also builds the replicas).   gpurun --gpus 8 --timeout 600 -- 'bash tools/dropin_scaling.sh'   (GPUS="1 8" for a subset)
set -u
mkdir -p gpurun_out
B=build/pathtracer_b200; R=oracle/_ref/resources
N=$(nvidia-smi -L | wc -l)
echo "{" > gpurun_out/r02_dropin_scaling.json
first=1
for g in ${GPUS:-1 2 4 8}; do
  [ $g -le $N ] || continue
  for split in tile sample; do
    [ $g -eq 1 ] && [ $split = sample ] && continue
    t=$(timeout 300 $B --render --scene gallery-materials --resources-root $R --size 1920 1080 --samples ${SPP:-256} --depth 8 --seed 1 --gpus $g --split $split --frames 3 --output gpurun_out/dropin_${g}_${split}.png 2>&1 | grep "Rendering took" | tail -1 | sed -E 's/.*took ([0-9.e-]+)s.*/\1/')
    [ $first -eq 1 ] || echo "," >> gpurun_out/r02_dropin_scaling.json; first=0
    echo "\"gpus${g}_${split}\": {\"seconds\": $t, \"spp_per_s\": $(python -c "print(${SPP:-256}/$t)")}" | tee -a gpurun_out/r02_dropin_scaling.json
  done
done
echo "}" >> gpurun_out/r02_dropin_scaling.json
cmp gpurun_out/dropin_1_tile.png gpurun_out/dropin_${N}_tile.png && echo "tile split over $N GPUs: PNG identical to 1 GPU"
