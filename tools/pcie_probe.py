"""PCIe floor of the host-buffer path: pinned H2D / D2H of the e2e payload sizes."""
import torch, time
dev = torch.device("cuda:0")
sizes = [1_600_000, 1_600_000, 1_600_000, 400_000, 1_200_000]   # pos, vel, forces, ids, images (bytes)
host = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
devb = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
big_h = torch.empty(sum(sizes), dtype=torch.uint8).pin_memory(); big_d = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
s0 = torch.cuda.Stream(); streams = [torch.cuda.Stream() for _ in sizes]
def timeit(fn, n=200):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e6
def five_one_stream():
    with torch.cuda.stream(s0):
        for h, d in zip(host, devb): d.copy_(h, non_blocking=True)
    s0.synchronize()
def five_streams():
    for h, d, st in zip(host, devb, streams):
        with torch.cuda.stream(st): d.copy_(h, non_blocking=True)
    for st in streams: st.synchronize()
def one_big():
    with torch.cuda.stream(s0): big_d.copy_(big_h, non_blocking=True)
    s0.synchronize()
def d2h():
    with torch.cuda.stream(s0): host[2].copy_(devb[2], non_blocking=True)
    s0.synchronize()
def both_dirs():
    with torch.cuda.stream(s0): big_d.copy_(big_h, non_blocking=True)
    with torch.cuda.stream(streams[0]): host[2].copy_(devb[2], non_blocking=True)
    s0.synchronize(); streams[0].synchronize()
for name, fn in (("H2D 6.4 MB, 5 copies, one stream + sync", five_one_stream), ("H2D 6.4 MB, 5 copies, 5 streams", five_streams),
                 ("H2D 6.4 MB, one copy", one_big), ("D2H 1.6 MB + sync", d2h), ("H2D 6.4 MB || D2H 1.6 MB", both_dirs)):
    us = timeit(fn)
    print(f"{name:45s} {us:7.1f} us")
