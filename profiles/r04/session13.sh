#!/bin/bash
# NOTE: the switch this session drives existed in the working tree for the experiment only (result: assoc_experiments.txt / e2e_variants.txt)
# where the association kernel's time goes at batch size (timing experiment, results wrong when set):
# FORMGPU_DEBUG_ASSOC_MODE=1 stops after the query's own unit, =2 after the centre probe
OUT=gpurun_out/r4o
mkdir -p $OUT
for v in 1 2; do
  FORMGPU_DEBUG_ASSOC_MODE=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-profile > $OUT/profile_mode$v.json 2> $OUT/profile_mode$v.err
  echo "mode $v: $(tail -1 $OUT/profile_mode$v.json)"
done
