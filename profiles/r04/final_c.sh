#!/bin/bash
# one GPU: the final bench line (both arms) and the other sensor shapes (configs[0], configs[2])
OUT=gpurun_out/r4x
mkdir -p $OUT
python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 500 $OUT/bench.json
python bench.py --sensor os1-64 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_os1-64.json 2> $OUT/bench_os1-64.err; echo "os1-64 rc=$?"
python bench.py --sensor vlp-16 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_vlp-16.json 2> $OUT/bench_vlp-16.err; echo "vlp-16 rc=$?"
python - <<'PY'
import json
for n in ("bench", "bench_os1-64", "bench_vlp-16"):
    try:
        d = json.loads(open(f"gpurun_out/r4x/{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["e2e"]["value"], d["single_sequence"]["value"], (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
