#!/bin/bash
# 256-bit loads of the map points: parity + times
OUT=gpurun_out/r4q
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-profile > $OUT/profile.json 2> $OUT/profile.err
echo "profile: $(tail -1 $OUT/profile.json)"
