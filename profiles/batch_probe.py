"""Development probe: ONE batch of M sequences replayed in lock step (device-resident scans),
the target of the ncu captures of the batched kernels.  Not a bench line.
usage: python profiles/batch_probe.py [M] [warmup] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from form_b200 import _capi, synth
from form_b200.pipeline import BatchReplay, Estimator

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W = int(sys.argv[2]) if len(sys.argv) > 2 else 12
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sensor = "os0-128"
rows, cols = synth.shape(sensor)
p = _capi.default_est_params(rows, cols, record_trace=1)
S = W + K
ests, ptrs, keep = [], [], []
for m in range(M):
    scans = [synth.scan(sensor, m, k) for k in range(S)]
    e = Estimator(p)
    for s in scans:
        e.register_scan(s)
    dv = [torch.from_numpy(s.view(np.uint8)).cuda() for s in scans]
    ests.append(e)
    keep.append(dv)
    ptrs.append([d.data_ptr() for d in dv])
torch.cuda.synchronize()
br = BatchReplay([e.trace() for e in ests], p)
br.run(0, W, ptrs)
torch.cuda.synchronize()
t0 = time.time()
t, rounds = br.run(W, S, ptrs)
torch.cuda.synchronize()
print(f"M={M} scans/s={M * K / t:.1f} submits={rounds} launches={br.launch_count()}")
