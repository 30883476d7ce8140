#!/bin/bash
# NOTE: the switch this session drives existed in the working tree for the experiment only (result: assoc_experiments.txt / e2e_variants.txt)
# cell search with the CTA's queries redistributed between the phases: parity + times
OUT=gpurun_out/r4p
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-profile > $OUT/profile.json 2> $OUT/profile.err
echo "profile: $(tail -1 $OUT/profile.json)"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-value > $OUT/value.json 2> $OUT/value.err
echo "value: $(tail -1 $OUT/value.json)"
