#!/usr/bin/env python
"""HBM ceilings for the traffic mixes of this path: pure write, 50/50 copy, and the 25 % read / 75 % write mix
of fk (one input array read, three times as many bytes written), with plain torch ops -- a yardstick for how
much of the gap between fk and the measured copy peak is the mix itself."""
import json

import torch

dev = torch.device("cuda", 0)
n = 256 * 1024 * 1024  # floats: 1 GiB
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
c = torch.empty(3 * n // 4 * 4 // 4 * 1, device=dev)  # scratch


def timed(fn, bytes_moved, steps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return bytes_moved / (e0.elapsed_time(e1) / steps * 1e-3) / 1e9


out = {
    "write_only_fill": timed(lambda: a.fill_(1.0), 4 * n),
    "write_only_zero": timed(lambda: a.zero_(), 4 * n),
    "copy_50_50": timed(lambda: b.copy_(a), 8 * n),
}
# 25 / 75: read n/4 floats, write 3n/4 floats (repeat = one read feeding three writes)
src = a[: n // 4]
dst = b[: 3 * n // 4].view(3, n // 4)
out["read25_write75_expand"] = timed(lambda: dst.copy_(src.expand(3, n // 4)), 4 * n)
out["read_only_sum"] = timed(lambda: a.sum(), 4 * n)
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
