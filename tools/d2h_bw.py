"""Pinned device->host bandwidth of the box (what bounds the e2e leg): one 2 GiB cudaMemcpyAsync, and 1024 x 2 MiB."""
import time
import torch
n = 1 << 28
d = torch.empty(n, dtype=torch.float64, device='cuda')
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
for tag, chunks in (('1 x 2 GiB', 1), ('1024 x 2 MiB', 1024)):
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        step = n // chunks
        for c in range(chunks):
            h[c * step:(c + 1) * step].copy_(d[c * step:(c + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print('D2H %s: %.1f GB/s' % (tag, n * 8 / dt / 1e9))
