#!/bin/bash
# two GPUs: the bench line under torchrun (independent sequences, weak scaling) and the 2-GPU tests
OUT=gpurun_out/r4y
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"; tail -c 700 $OUT/bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $OUT/bench_reference_n2.json 2> $OUT/bench_reference_n2.err; echo "reference n2 rc=$?"; tail -c 300 $OUT/bench_reference_n2.json
python -m pytest tests/test_gpu_sharded.py tests/test_gpu_batch.py -m gpu -q > $OUT/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_2gpu.log
tail -3 $OUT/pytest_2gpu.log
