#!/bin/bash
# parity run + bench after: restructured cached evaluation (6 barriers, host-resolved entry pointer,
# parallel basis fill), scan prefetch of the replay driver; A/B of the prefetch on the e2e leg
OUT=gpurun_out/r4e
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
run() { name=$1; shift; "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(tail -1 $OUT/$name.json | cut -c1-300)"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run bench $B
FORM_REPLAY_PREFETCH=0 run e2e_noprefetch $B --only-e2e
run e2e_prefetch $B --only-e2e
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r4e/bench.json").read().strip().splitlines()[-1])
kg = d["roofline"]["kernel_groups"]
print("bench", d["value"], d["e2e"]["value"], {k: round(v["ms_per_scan"] * 1e3, 1) for k, v in kg.items()})
print("  single", d["single_sequence"]["value"], d["single_sequence"]["e2e"], d["single_sequence"]["kernel_ms_per_scan"])
PY
