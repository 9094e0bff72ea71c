"""Pinned device->host bandwidth of the box (what bounds the e2e leg): one device alone, then all visible devices at once."""
import time
import torch
ndev = torch.cuda.device_count()
n = 1 << 29   # 4 GiB per device
host = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(ndev)]
dev = [torch.empty(n, dtype=torch.float64, device='cuda:%d' % g) for g in range(ndev)]
streams = [torch.cuda.Stream(device=g) for g in range(ndev)]


def run(gs):
    for g in gs:
        torch.cuda.synchronize(g)
    t = time.perf_counter()
    for g in gs:
        with torch.cuda.stream(streams[g]):
            host[g].copy_(dev[g], non_blocking=True)
    for g in gs:
        streams[g].synchronize()
    return time.perf_counter() - t


for gs in [[0]] + ([list(range(ndev))] if ndev > 1 else []):
    run(gs)
    dt = run(gs)
    print('D2H, devices %s at once, 4 GiB each into pinned memory: %.1f GB/s total' % (gs, len(gs) * n * 8 / dt / 1e9), flush=True)
