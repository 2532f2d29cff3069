"""Does pinned-memory placement (NUMA node) explain the spread of the e2e number?  Prints the topology the
container can see and the H2D rate of a 21 MB pinned buffer allocated under each node's CPU affinity."""
import glob
import os
import subprocess

import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def main():
    allowed = sorted(os.sched_getaffinity(0))
    print("allowed cpus:", len(allowed), allowed[:4], "...", allowed[-4:])
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        nodes[int(d.rsplit("node", 1)[1])] = cpulist(open(d + "/cpulist").read())
    for n, c in nodes.items():
        print(f"node {n}: {len(c)} cpus, {len(set(c) & set(allowed))} allowed")
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"],
                           capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
        for line in q:
            idx, bus = [t.strip() for t in line.split(",")]
            bus = bus.lower()
            if bus.startswith("00000000:"):
                bus = "0000:" + bus.split(":", 1)[1]
            p = f"/sys/bus/pci/devices/{bus}/numa_node"
            print("gpu", idx, bus, "numa_node", open(p).read().strip() if os.path.exists(p) else "?")
    except Exception as e:  # noqa: BLE001
        print("nvidia-smi query failed:", e)
    dev = torch.device("cuda", 0)
    dst = torch.empty(21 * 1024 * 1024 // 4, device=dev)
    for n, cpus in list(nodes.items()) + [(-1, allowed)]:
        use = sorted(set(cpus) & set(allowed))
        if not use:
            continue
        os.sched_setaffinity(0, use)
        src = torch.empty(dst.numel()).pin_memory()
        src.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        print(f"pinned under node {n} affinity ({len(use)} cpus): {10 * dst.numel() * 4 / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s")
        os.sched_setaffinity(0, allowed)
        del src


if __name__ == "__main__":
    main()
