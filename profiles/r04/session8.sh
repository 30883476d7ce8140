#!/bin/bash
OUT=gpurun_out/r4j
mkdir -p $OUT
run() { name=$1; shift; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(tail -1 $OUT/$name.json | cut -c1-200)"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
FORMGPU_PACK_DMA=0 run e2e_b12_zerocopy $B --only-e2e --batches-per-gpu 12
FORMGPU_PACK_DMA=0 FORM_REPLAY_PREFETCH=0 run e2e_b12_zerocopy_noprefetch $B --only-e2e --batches-per-gpu 12
FORM_REPLAY_PREFETCH=0 run e2e_b12_noprefetch $B --only-e2e --batches-per-gpu 12
run e2e_b12 $B --only-e2e --batches-per-gpu 12
