#!/bin/bash
# one GPU: parity of the last changes and the other sensor shapes (configs[0], configs[2])
OUT=gpurun_out/r4x
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_1gpu.log
tail -3 $OUT/pytest_1gpu.log
python bench.py --sensor os1-64 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_os1-64.json 2> $OUT/bench_os1-64.err; echo "os1-64 rc=$?"
python bench.py --sensor vlp-16 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_vlp-16.json 2> $OUT/bench_vlp-16.err; echo "vlp-16 rc=$?"
python - <<'PY'
import json
for n in ("bench_os1-64", "bench_vlp-16"):
    try:
        d = json.loads(open(f"gpurun_out/r4x/{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["e2e"]["value"], d["single_sequence"]["value"], d["config"]["sequences_per_gpu"], d["roofline"]["kernel"], d["roofline"]["frac"])
    except Exception as e:
        print(n, "failed", e)
PY
