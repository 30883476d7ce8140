"""Development probe for the e2e gap: ONE batch of M sequences replayed twice -
scans resident in HBM, then scans in page-locked host memory with f64 keypoints written back -
with CUDA events around every launch, so the kernel groups that stretch in the host-scan
variant (extract_pack writes over PCIe; extract_select waits for the side-stream upload) show up
next to the wall time of each variant.  Not a bench line.
usage: python profiles/e2e_probe.py [M] [warmup] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from form_b200 import _capi, synth
from form_b200.pipeline import BatchReplay, Estimator

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W = int(sys.argv[2]) if len(sys.argv) > 2 else 12
K = int(sys.argv[3]) if len(sys.argv) > 3 else 6
sensor = "os0-128"
rows, cols = synth.shape(sensor)
n_points = rows * cols
p = _capi.default_est_params(rows, cols, record_trace=1)
S = W + K
ests, dev_ptrs, host_ptrs, keep = [], [], [], []
for m in range(M):
    pinned = torch.empty((S, n_points, 4), dtype=torch.float32).pin_memory()
    host = pinned.numpy().view(_capi.POINT4F).reshape(S, n_points)
    for k in range(S):
        host[k][:] = synth.scan(sensor, m, k, 1)
    e = Estimator(p)
    for k in range(S):
        e.register_scan(host[k])
    dev = pinned.cuda()
    ests.append(e)
    keep.append((pinned, host, dev))
    dev_ptrs.append([dev[k].data_ptr() for k in range(S)])
    host_ptrs.append([host[k].ctypes.data for k in range(S)])
torch.cuda.synchronize()
traces = [e.trace() for e in ests]
for name, ptrs, on_device in (("device scans", dev_ptrs, True), ("host scans (e2e)", host_ptrs, False)):
    br = BatchReplay(traces, p)
    br.run(0, W, ptrs, on_device=on_device)
    torch.cuda.synchronize()
    br.profile_read()
    br.profile_enable(True)
    t0 = time.time()
    t, rounds = br.run(W, S, ptrs, on_device=on_device)
    torch.cuda.synchronize()
    wall = time.time() - t0
    prof = br.profile_read()
    br.profile_enable(False)
    print(f"== {name}: {M * K / t:.1f} scans/s under per-launch events, {rounds} submits, wall {wall * 1e3:.1f} ms")
    for g, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        if v["launches"]:
            print(f"   {g:16s} {1e3 * v['ms'] / (M * K):8.1f} us/scan  {v['launches']:6d} launches "
                  f"{1e3 * v['ms'] / v['launches']:8.1f} us/launch")
    br.close()
