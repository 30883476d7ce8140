#!/bin/bash
# NOTE: the switch this session drives existed in the working tree for the experiment only (result: assoc_experiments.txt / e2e_variants.txt)
OUT=gpurun_out/r4k
mkdir -p $OUT
run() { name=$1; shift; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(tail -1 $OUT/$name.json | cut -c1-200)"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
FORMGPU_SM_UPLOAD=8 run e2e_sm8 $B --only-e2e
FORMGPU_SM_UPLOAD=32 run e2e_sm32 $B --only-e2e
FORMGPU_SM_UPLOAD=32 FORM_REPLAY_PREFETCH=0 run e2e_sm32_noprefetch $B --only-e2e
run e2e_dma $B --only-e2e
run value $B --only-value
