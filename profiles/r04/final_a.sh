#!/bin/bash
# final evidence of the round on one GPU: parity, the bench line (both arms), the ncu launch list of
# a reduced bench command, and one `ncu --set full` capture of the batched kernels
OUT=gpurun_out/r4z
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_1gpu.log
tail -3 $OUT/pytest_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 600 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?"; tail -c 400 $OUT/bench_reference.json
# launch list: the same reduced command first without ncu
R="python bench.py --steps 4 --warmup 3 --preroll 12 --sequences-per-gpu 16 --batches-per-gpu 2 --no-cpu-baseline --only-value"
$R > $OUT/reduced_plain.json 2> $OUT/reduced_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'batch|normals_rows|eval_global' --launch-skip 600 -c 2500 --csv --log-file $OUT/launches.csv $R > $OUT/reduced_ncu.log 2>&1
echo "launch list rows: $(wc -l < $OUT/launches.csv)"
python profiles/batch_probe.py 8 10 3 > $OUT/probe_plain.log 2>&1 && \
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'batch|normals_rows|eval_global' --launch-skip 520 -c 30 -f -o /tmp/full python profiles/batch_probe.py 8 10 3 > $OUT/ncu_full.log 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > $OUT/full_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/full.ncu-rep); echo "report bytes $sz"
if [ $sz -lt 42000000 ]; then cp /tmp/full.ncu-rep $OUT/full.ncu-rep; fi
du -sh $OUT
