#!/bin/bash
# `ncu --set full` capture of the batched kernels of profiles/batch_probe.py (8 sequences); the
# report's raw / source pages are exported on the box (the .ncu-rep itself is kept only if small)
set -x
OUT=gpurun_out/r4b
mkdir -p $OUT
python profiles/batch_probe.py 8 10 3 > $OUT/probe_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'batch|normals_rows|eval_global' --launch-skip 520 -c 30 -f -o /tmp/full \
  python profiles/batch_probe.py 8 10 3 > $OUT/ncu_full.log 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i /tmp/full.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
gzip -9 $OUT/source.csv
ls -la /tmp/full.ncu-rep $OUT
sz=$(stat -c %s /tmp/full.ncu-rep)
if [ $sz -lt 45000000 ]; then cp /tmp/full.ncu-rep $OUT/full.ncu-rep; fi
du -sh $OUT
