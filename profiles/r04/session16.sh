#!/bin/bash
# speculative sector-parallel walks in the select kernel: parity of every extraction test, then the
# kernel-group times with 128 / 192 threads per row
OUT=gpurun_out/r4r
mkdir -p $OUT
python -m pytest tests/test_gpu_extract.py tests/test_gpu_many_rows.py tests/test_gpu_reference.py tests/test_gpu_batch.py -m gpu -x -q > $OUT/pytest_extract.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_extract.log
tail -4 $OUT/pytest_extract.log
for t in 128 192; do
  FORMGPU_SELECT_THREADS=$t python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-profile > $OUT/profile_t$t.json 2> $OUT/profile_t$t.err
  echo "threads $t: $(tail -1 $OUT/profile_t$t.json)"
done
