#!/bin/bash
# NOTE: the switch this session drives existed in the working tree for the experiment only (result: assoc_experiments.txt / e2e_variants.txt)
# association kernel variants (FORMGPU_ASSOC_VARIANT): occupancy bound / bucket points in flight;
# per-kernel-group CUDA-event times of the one-batch profile replay, us per scan
OUT=gpurun_out/r4n
mkdir -p $OUT
for v in 0 1 2 3; do
  FORMGPU_ASSOC_VARIANT=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-profile > $OUT/profile_v$v.json 2> $OUT/profile_v$v.err
  echo "variant $v: $(tail -1 $OUT/profile_v$v.json)"
done
