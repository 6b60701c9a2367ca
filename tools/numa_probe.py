"""Does the host-buffer path depend on which NUMA node the pinned memory lives on?  H2D / D2H rate per node."""
import glob, os, time
import torch

dev = torch.device("cuda:0")
props = torch.cuda.get_device_properties(0)
bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
def rd(p):
    try:
        return open(p).read().strip()
    except OSError as e:
        return f"? ({e.__class__.__name__})"
print("GPU", props.name, "pci", bus, "numa_node", rd(f"/sys/bus/pci/devices/{bus}/numa_node"), "local_cpulist", rd(f"/sys/bus/pci/devices/{bus}/local_cpulist"))
print("process affinity", sorted(os.sched_getaffinity(0))[:8], "... n =", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
def parse(cl):
    out = []
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out
allowed = os.sched_getaffinity(0)
d = torch.empty(6_400_000, dtype=torch.uint8, device=dev)
def rate(h, n=50):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5): d.copy_(h, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(n): d.copy_(h, non_blocking=True)
        e1.record(s); s.synchronize()
        up = 6.4e6 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
        e0.record(s)
        for _ in range(n): h.copy_(d, non_blocking=True)
        e1.record(s); s.synchronize()
        down = 6.4e6 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return up, down
h = torch.empty(6_400_000, dtype=torch.uint8).pin_memory(); h.fill_(1)
print("default placement: H2D %.1f GB/s, D2H %.1f GB/s" % rate(h))
for nd in nodes:
    cpus = [c for c in parse(rd(nd + "/cpulist")) if c in allowed]
    if not cpus:
        print(os.path.basename(nd), "no allowed cpus"); continue
    os.sched_setaffinity(0, cpus)
    time.sleep(0.01)
    h = torch.empty(6_400_000, dtype=torch.uint8)
    h.fill_(1)            # first touch on this node
    h = h.pin_memory(); h.fill_(2)
    up, down = rate(h)
    print(f"{os.path.basename(nd)} ({len(cpus)} cpus): H2D {up:.1f} GB/s, D2H {down:.1f} GB/s")
os.sched_setaffinity(0, allowed)
