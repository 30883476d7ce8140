#!/bin/bash
# last run of the round: full parity suite and the bench line with the speculative select walks
OUT=gpurun_out/r4s
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_1gpu.log
tail -3 $OUT/pytest_1gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r4s/bench.json").read().strip().splitlines()[-1])
kg = d["roofline"]["kernel_groups"]
print("bench", d["value"], d["e2e"]["value"], {k: round(v["ms_per_scan"] * 1e3, 1) for k, v in kg.items()})
print("  single", d["single_sequence"]["value"], d["single_sequence"]["e2e"], d["single_sequence"]["kernel_ms_per_scan"])
print("  roofline", d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["dram_frac"])
PY
