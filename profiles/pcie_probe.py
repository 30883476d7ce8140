"""Host<->device copy bandwidth of the box (page-locked memory), the ceiling of the e2e leg:
every scan is 2 MiB up, ~1.5 MB of keypoints + blocks down.  usage: python profiles/pcie_probe.py"""
import time

import torch

torch.cuda.init()
dev = torch.device("cuda", 0)


def bw(nbytes, direction, streams=1, reps=40):
    hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(streams)]
    ds = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    for warm in (True, False):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3 if warm else reps):
            for h, d, s in zip(hs, ds, ss):
                with torch.cuda.stream(s):
                    if direction == "h2d":
                        d.copy_(h, non_blocking=True)
                    elif direction == "d2h":
                        h.copy_(d, non_blocking=True)
                    else:
                        d.copy_(h, non_blocking=True)
                        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    mult = 2 if direction == "both" else 1
    return mult * streams * reps * nbytes / dt / 1e9


for size in (2 << 20, 64 << 20):
    for direction in ("h2d", "d2h", "both"):
        for streams in (1, 4):
            print(f"{size >> 20:3d} MiB {direction:5s} x{streams}: {bw(size, direction, streams):6.1f} GB/s")
