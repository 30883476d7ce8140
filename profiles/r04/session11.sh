#!/bin/bash
# sensitivity of the step time to the work of one kernel: the association / select kernel is
# launched twice (the extra association launch is a dry run: whole search, no side effects)
OUT=gpurun_out/r4m
mkdir -p $OUT
run() { name=$1; shift; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(tail -1 $OUT/$name.json | cut -c1-200)"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run value_base $B --only-value
FORMGPU_DEBUG_ASSOC_REPEAT=1 run value_assoc_x2 $B --only-value
FORMGPU_DEBUG_SELECT_REPEAT=1 run value_select_x2 $B --only-value
run value_base2 $B --only-value
