"""Pinned device->host bandwidth of the box (what bounds the e2e leg): small and large pinned destinations."""
import sys
import time
import torch
for gib in (2, 16, 64):
    n = gib << 27
    try:
        h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    except Exception as e:
        print('pin %d GiB failed: %s' % (gib, e))
        break
    d = torch.empty(min(n, 1 << 30), dtype=torch.float64, device='cuda')
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for off in range(0, n, d.numel()):
            h[off:off + d.numel()].copy_(d[:min(d.numel(), n - off)], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print('D2H into a %d GiB pinned buffer (8 GiB pieces): %.1f GB/s' % (gib, n * 8 / dt / 1e9), flush=True)
    del h
