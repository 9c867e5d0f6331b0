"""Host topology probe: NUMA node of every visible GPU, the CPUs this process may run on, whether the
memory policy of the calling thread can be set (set_mempolicy), and pinned-copy bandwidth to GPU 0 from
memory bound to each node."""
import ctypes, glob, os, sys
import torch

print("cpus allowed:", sorted(os.sched_getaffinity(0)))
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("numa nodes:", nodes)
for n in nodes:
    try:
        print(" node", n, "cpus", open(f"/sys/devices/system/node/node{n}/cpulist").read().strip())
    except OSError as e:
        print(" node", n, e)
import pynvml
pynvml.nvmlInit()
for i in range(torch.cuda.device_count()):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
    bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
    path = "/sys/bus/pci/devices/" + bdf.lower()[-12:] + "/numa_node"
    try:
        nn = open(path).read().strip()
    except OSError as e:
        nn = str(e)
    print("gpu", i, bdf, "numa_node", nn)
libc = ctypes.CDLL("libc.so.6", use_errno=True)
MPOL_DEFAULT, MPOL_BIND, MPOL_PREFERRED = 0, 2, 1
def set_policy(mode, node):
    mask = ctypes.c_ulong(0 if node is None else 1 << node)
    r = libc.syscall(238, mode, ctypes.byref(mask), ctypes.c_ulong(64))
    return r, ctypes.get_errno()
dev = torch.device("cuda", 0)
n = 256 * 1000 * 1000
dst = torch.empty(n, dtype=torch.uint8, device=dev)
for node in [None] + nodes:
    r = set_policy(MPOL_DEFAULT if node is None else MPOL_BIND, node)
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    set_policy(MPOL_DEFAULT, None)
    for _ in range(2): dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dst.copy_(host, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("policy node", node, "set_mempolicy ->", r, " H2D to gpu0: %.1f GB/s" % (n * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
    del host
