#!/bin/bash
# parity run + bench after: near-octant-first cell search, convergent chunk evaluation in the
# thread-per-pick normals, pipelined loads in the moment kernel / segment prefix; batch-count sweep
set -x
OUT=gpurun_out/r4c
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-value --batches-per-gpu 32 > $OUT/bench_b32.json 2> $OUT/bench_b32.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-value --batches-per-gpu 8 > $OUT/bench_b8.json 2> $OUT/bench_b8.err
python - <<'PY'
import json
for n in ("bench", "bench_b32", "bench_b8"):
    try:
        d = json.loads(open(f"gpurun_out/r4c/{n}.json").read().strip().splitlines()[-1])
        kg = d.get("roofline", {}).get("kernel_groups", {})
        print(n, d["value"], d.get("e2e", {}).get("value"), {k: round(v["ms_per_scan"] * 1e3, 1) for k, v in kg.items()})
        if "single_sequence" in d: print("  single", d["single_sequence"]["value"], d["single_sequence"]["kernel_ms_per_scan"])
    except Exception as e:
        print(n, "failed", e)
PY
