#!/bin/bash
# e2e gap probe (host scans with / without keypoint outputs) and batch-count sweep
OUT=gpurun_out/r4d
mkdir -p $OUT
run() { name=$1; shift; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(tail -1 $OUT/$name.json)"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run e2e_b16 $B --only-e2e
FORM_REPLAY_SKIP_KEYPOINTS=1 run e2e_b16_nokp $B --only-e2e
run e2e_b8 $B --only-e2e --batches-per-gpu 8
run value_b4 $B --only-value --batches-per-gpu 4
run value_b6 $B --only-value --batches-per-gpu 6
run value_b8 $B --only-value --batches-per-gpu 8
run value_b12 $B --only-value --batches-per-gpu 12
