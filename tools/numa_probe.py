"""Host topology of a multi-GPU box and what page placement does to the aggregate H2D/D2H rate.

Prints the NUMA layout (nodes, CPUs, the node each GPU hangs off) and then times concurrent pinned copies on every
visible GPU with the pinned buffers placed (a) wherever the default policy puts them, (b) on the GPU's own node
(set_mempolicy(MPOL_BIND) around the allocation), (c) interleaved over all nodes.  One thread per GPU.

    python tools/numa_probe.py [MB per copy] [copies]
"""
import ctypes
import glob
import json
import os
import subprocess
import sys
import threading
import time

import torch

MPOL_DEFAULT, MPOL_BIND, MPOL_INTERLEAVE = 0, 2, 3
_libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238  # x86_64


def set_mempolicy(mode, nodes):
    mask = 0
    for n in nodes:
        mask |= 1 << n
    m = ctypes.c_ulong(mask)
    rc = _libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(m) if nodes else None, 64 if nodes else 0)
    return rc, ctypes.get_errno()


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def topology():
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    info = {"nodes": {}, "gpus": []}
    for n in nodes:
        cpus = open(f"/sys/devices/system/node/node{n}/cpulist").read().strip()
        mem = open(f"/sys/devices/system/node/node{n}/meminfo").read().split("MemTotal:")[1].split()[0]
        info["nodes"][n] = {"cpus": cpus, "mem_gb": int(mem) / 1e6}
    for i in range(torch.cuda.device_count()):
        bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
        info["gpus"].append({"index": i, "bus": bus})
    q = sh("nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader")
    for line in q.strip().splitlines():
        idx, bus = [x.strip() for x in line.split(",")]
        bus = bus.lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        path = f"/sys/bus/pci/devices/{bus}/numa_node"
        node = open(path).read().strip() if os.path.exists(path) else "?"
        cpul = f"/sys/bus/pci/devices/{bus}/local_cpulist"
        info["gpus"][int(idx)].update({"bus": bus, "numa_node": node, "local_cpulist": open(cpul).read().strip() if os.path.exists(cpul) else "?"})
    info["allowed_cpus"] = sorted(os.sched_getaffinity(0))
    st = open("/proc/self/status").read()
    info["mems_allowed_list"] = st.split("Mems_allowed_list:")[1].split()[0]
    return info


def run(policy, info, mb, copies, direction):
    n = torch.cuda.device_count()
    nodes = sorted(info["nodes"])
    bufs, devs, streams = [], [], []
    for i in range(n):
        node = info["gpus"][i].get("numa_node", "?")
        if policy == "local" and node not in ("?", "-1"):
            set_mempolicy(MPOL_BIND, [int(node)])
        elif policy == "interleave":
            set_mempolicy(MPOL_INTERLEAVE, nodes)
        elif policy.startswith("node"):
            set_mempolicy(MPOL_BIND, [int(policy[4:])])
        h = torch.empty(mb << 20, dtype=torch.uint8)
        h.fill_(i + 1)  # first touch under the policy
        h = h.pin_memory()
        set_mempolicy(MPOL_DEFAULT, [])
        bufs.append(h)
        with torch.cuda.device(i):
            devs.append(torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{i}"))
            streams.append((torch.cuda.Stream(device=i), torch.cuda.Stream(device=i)))
    for i in range(n):
        torch.cuda.synchronize(i)
    barrier = threading.Barrier(n + 1)

    def worker(i):
        with torch.cuda.device(i):
            s_up, s_dn = streams[i]
            barrier.wait()
            for _ in range(copies):
                if direction in ("h2d", "both"):
                    with torch.cuda.stream(s_up):
                        devs[i].copy_(bufs[i], non_blocking=True)
                if direction in ("d2h", "both"):
                    with torch.cuda.stream(s_dn):
                        bufs[i][: (mb << 20) // 4].copy_(devs[i][: (mb << 20) // 4], non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
            barrier.wait()

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
    for t in ts:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in ts:
        t.join()
    up = n * copies * (mb << 20) / dt / 1e9 if direction in ("h2d", "both") else 0.0
    dn = n * copies * ((mb << 20) // 4) / dt / 1e9 if direction in ("d2h", "both") else 0.0
    return {"policy": policy, "dir": direction, "h2d_GBps": round(up, 1), "d2h_GBps": round(dn, 1), "s": round(dt, 3)}


def host_stream_bw(threads):
    """Aggregate host read bandwidth (numpy sum over private 256 MB arrays, one per thread)."""
    import numpy as np

    arrs = [np.ones(256 << 20, dtype=np.uint8) for _ in range(threads)]
    barrier = threading.Barrier(threads + 1)

    def w(a):
        barrier.wait()
        for _ in range(4):
            a.view(np.uint64).sum()
        barrier.wait()

    ts = [threading.Thread(target=w, args=(a,)) for a in arrs]
    for t in ts:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in ts:
        t.join()
    return threads * 4 * (256 << 20) / dt / 1e9


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    copies = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    print(sh("lscpu | egrep -i 'model name|socket|numa|^cpu\\(s\\)|thread'"))
    print(sh("nvidia-smi topo -m"))
    print(sh("numactl -H 2>&1 | head -20"))
    info = topology()
    print(json.dumps(info, indent=1))
    ncpu = len(info["allowed_cpus"])
    print("host read bandwidth, %d threads: %.1f GB/s" % (ncpu, host_stream_bw(ncpu)), flush=True)
    policies = ["default", "local", "interleave"] + ([f"node{n}" for n in sorted(info["nodes"])] if len(info["nodes"]) > 1 else [])
    for policy in policies:
        for direction in ("h2d", "both"):
            print(json.dumps(run(policy, info, mb, copies, direction)), flush=True)


if __name__ == "__main__":
    main()
