#!/bin/bash
# host-side round timing of the batch engine (FORMGPU_BATCH_TRACE=1): value leg vs e2e leg
OUT=gpurun_out/r4f
mkdir -p $OUT
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
FORMGPU_BATCH_TRACE=1 $B --only-value > $OUT/value.json 2> $OUT/value.err; tail -1 $OUT/value.json
FORMGPU_BATCH_TRACE=1 $B --only-e2e > $OUT/e2e.json 2> $OUT/e2e.err; tail -1 $OUT/e2e.json
grep -h "formgpu batch" $OUT/value.err | head -9
echo ----
grep -h "formgpu batch" $OUT/e2e.err | head -9
