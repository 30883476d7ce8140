#!/bin/bash
# round-2 session: block-DMA validation (pytest -m gpu), A/B of the batched block transfer, and one
# `ncu --set full` capture of every batched kernel of profiles/batch_probe.py (8 sequences)
set -x
OUT=gpurun_out/r4a
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_dma1.json 2> $OUT/bench_dma1.err
FORMGPU_BLOCK_DMA=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --only-value > $OUT/bench_dma0.json 2> $OUT/bench_dma0.err
python - <<'PY'
import json
for n in ("dma1", "dma0"):
    try:
        d = json.loads(open(f"gpurun_out/r4a/bench_{n}.json").read().strip().splitlines()[-1])
        kg = d.get("roofline", {}).get("kernel_groups", {})
        print(n, d["value"], d.get("e2e", {}).get("value"), {k: round(v["ms_per_scan"] * 1e3, 1) for k, v in kg.items()})
        if "single_sequence" in d: print("  single", d["single_sequence"]["value"], d["single_sequence"]["kernel_ms_per_scan"])
    except Exception as e:
        print(n, "failed", e)
PY
python profiles/batch_probe.py 8 10 3 > $OUT/probe_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'batch|normals_rows|eval_global' --launch-skip 500 -c 48 -f -o $OUT/full \
  python profiles/batch_probe.py 8 10 3 > $OUT/ncu_full.log 2>&1
ls -la $OUT
