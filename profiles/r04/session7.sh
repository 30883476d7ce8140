#!/bin/bash
# parity + bench after the API-call diet of the batched submit (one completion word per submission,
# copies on the batch's stream, map requests staged with the arguments); batch-count sweep
OUT=gpurun_out/r4g
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
run() { name=$1; shift; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(tail -1 $OUT/$name.json | cut -c1-200)"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run bench $B
run value_b8 $B --only-value --batches-per-gpu 8
run value_b12 $B --only-value --batches-per-gpu 12
run e2e_b8 $B --only-e2e --batches-per-gpu 8
run e2e_b12 $B --only-e2e --batches-per-gpu 12
FORMGPU_NO_STREAM_MEMOPS=1 run value_b16_flagkernel $B --only-value
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r4g/bench.json").read().strip().splitlines()[-1])
kg = d["roofline"]["kernel_groups"]
print("bench", d["value"], d["e2e"]["value"], d["gpu_launches"], {k: round(v["ms_per_scan"] * 1e3, 1) for k, v in kg.items()})
print("  single", d["single_sequence"]["value"], d["single_sequence"]["e2e"])
PY
