#!/bin/bash
# dynamic batch claiming in the replay driver: parity of the batch tests, thread / batch sweep
OUT=gpurun_out/r4l
mkdir -p $OUT
python -m pytest tests/test_gpu_batch.py tests/test_gpu_stages.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
run() { name=$1; shift; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(tail -1 $OUT/$name.json | cut -c1-200)"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run value_g12_t12 $B --only-value
run value_g12_t6 $B --only-value --host-threads 6
run value_g16_t8 $B --only-value --batches-per-gpu 16 --host-threads 8
run value_g8_t8 $B --only-value --batches-per-gpu 8
run value_g8_t4 $B --only-value --batches-per-gpu 8 --host-threads 4
run e2e_g12_t12 $B --only-e2e
run e2e_g12_t6 $B --only-e2e --host-threads 6
