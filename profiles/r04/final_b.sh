#!/bin/bash
# two GPUs: the bench line under torchrun (independent sequences, weak scaling)
OUT=gpurun_out/r4y
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"; tail -c 900 $OUT/bench_n2.json
