#!/usr/bin/env python
"""Pure-write and pure-read HBM bandwidth of this GPU (torch fill / sum over 4 GiB), next to the copy figure
MEASURED_PEAKS.json uses: the training chains are write-dominated, so this is the ceiling that applies to them."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, bytes_, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return bytes_ / (best * 1e-3) / 1e9
print("fill  (write only)  %.0f GB/s" % t(lambda: a.fill_(1.0), 4 * n))
print("sum   (read only)   %.0f GB/s" % t(lambda: a.sum(), 4 * n))
print("copy  (read+write)  %.0f GB/s" % t(lambda: b.copy_(a), 8 * n))
print("cudaMemset          %.0f GB/s" % t(lambda: a.zero_(), 4 * n))
