"""Host-side probe for the multi-GPU e2e legs (not product code): why do 8 ranks together reach only ~90 GB/s into host memory when one
reaches 52?  Prints the box topology (GPU <-> NUMA node, CPU lists) and measures the aggregate pinned D2H and H2D rate of k = 1, 2, 4, 8
GPUs copying at once, with the page-locked buffers allocated (a) by the main thread and (b) by a thread bound to the CPUs of the GPU's
NUMA node (first touch then lands in local memory)."""
import glob, os, subprocess, sys, threading, time
import torch

def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout
    except Exception as e:
        return "failed: %s\n" % e

print(sh("nvidia-smi topo -m"))
print("cpus", os.cpu_count(), "affinity", sorted(os.sched_getaffinity(0)))
for node in sorted(glob.glob("/sys/devices/system/node/node*")):
    try:
        print(os.path.basename(node), "cpus", open(node + "/cpulist").read().strip(), "|", open(node + "/meminfo").read().split("\n")[0].strip())
    except Exception as e:
        print(node, e)
ng = torch.cuda.device_count()
gpu_node = []
for d in range(ng):
    bdf = torch.cuda.get_device_properties(d).pci_bus_id if hasattr(torch.cuda.get_device_properties(d), "pci_bus_id") else None
    if bdf is None:
        q = sh("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i %d" % d).strip()
        bdf = q[-12:].lower() if q else ""
    node = -1
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf.lower()).read())
    except Exception:
        pass
    gpu_node.append(node)
    print("gpu", d, "pci", bdf, "numa_node", node)
SIZE = 1 << 30
REPS = 6
dev_bufs = [torch.empty(SIZE, dtype=torch.uint8, device="cuda:%d" % d) for d in range(ng)]
streams = [torch.cuda.Stream(device=d) for d in range(ng)]


def alloc_pinned(d, bind):
    out = [None]
    def work():
        if bind and gpu_node[d] >= 0:
            try:
                cl = open("/sys/devices/system/node/node%d/cpulist" % gpu_node[d]).read().strip()
                cpus = set()
                for part in cl.split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
            except Exception as e:
                print("bind failed", e)
        torch.cuda.set_device(d)
        t = torch.empty(SIZE, dtype=torch.uint8).pin_memory()
        t.fill_(1)
        out[0] = t
    th = threading.Thread(target=work); th.start(); th.join()
    return out[0]


for bind in (False, True):
    host = [alloc_pinned(d, bind) for d in range(ng)]
    for k in (1, 2, 4, 8):
        if k > ng:
            continue
        for direction in ("D2H", "H2D"):
            for d in range(k):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for _ in range(REPS):
                for d in range(k):
                    with torch.cuda.stream(streams[d]):
                        if direction == "D2H":
                            host[d].copy_(dev_bufs[d], non_blocking=True)
                        else:
                            dev_bufs[d].copy_(host[d], non_blocking=True)
            for d in range(k):
                streams[d].synchronize()
            dt = time.perf_counter() - t0
            print("pinned buffers %s: %d GPUs %s at once: %.1f GB/s aggregate (%.1f per GPU)" % (
                "allocated from a thread bound to the GPU's NUMA node" if bind else "allocated by the main thread", k, direction,
                k * REPS * SIZE / dt / 1e9, REPS * SIZE / dt / 1e9), flush=True)
    del host
