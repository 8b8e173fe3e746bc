#!/usr/bin/env python
"""Reference HBM bandwidths of this GPU with torch ops (CUDA events, best of 10 over 4 GiB):
write-only (fill), read-only (sum), copy (read + write).  Context for the write-bound chain kernels."""
import json
import torch
n = 1 << 30                                   # fp32 elements = 4 GiB
a = torch.empty(n, device="cuda")
b = torch.empty(n, device="cuda")
def best(fn, reps=10):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
gb = n * 4 / 1e9
res = {"write_only_gbs": gb / best(lambda: a.fill_(1.0)) * 1e3,
       "read_only_gbs": gb / best(lambda: a.sum()) * 1e3,
       "copy_gbs": 2 * gb / best(lambda: b.copy_(a)) * 1e3}
print(json.dumps(res))
