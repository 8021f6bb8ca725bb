#!/usr/bin/env python
"""Copy-engine behaviour on this box: H2D rate by transfer size on one stream, on two alternating streams, and with a
concurrent D2H stream; plus the same for D2H.  Pinned host memory, CUDA events."""
import torch

dev = torch.device("cuda", 0)
TOTAL = 1 << 30


def run(mb, n_streams, direction, other_dir_mb=0):
    nb = mb << 20
    reps = max(4, TOTAL // nb)
    hs = [torch.empty(nb, dtype=torch.uint8).pin_memory() for _ in range(n_streams)]
    ds = [torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(n_streams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    other = None
    if other_dir_mb:
        ob = other_dir_mb << 20
        oh, od, os_ = torch.empty(ob, dtype=torch.uint8).pin_memory(), torch.empty(ob, dtype=torch.uint8, device=dev), torch.cuda.Stream(device=dev)
        other = (oh, od, os_, max(4, TOTAL // ob))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ends = []
    e0.record(streams[0])
    for s in streams[1:]:
        s.wait_event(e0)
    if other:
        other[2].wait_event(e0)
        with torch.cuda.stream(other[2]):
            for _ in range(other[3]):
                (other[0].copy_(other[1], non_blocking=True) if direction == "h2d" else other[1].copy_(other[0], non_blocking=True))
    for i in range(reps):
        k = i % n_streams
        with torch.cuda.stream(streams[k]):
            (ds[k].copy_(hs[k], non_blocking=True) if direction == "h2d" else hs[k].copy_(ds[k], non_blocking=True))
    for s in streams[1:]:
        ev = torch.cuda.Event()
        ev.record(s)
        streams[0].wait_event(ev)
    e1.record(streams[0])
    torch.cuda.synchronize()
    return reps * nb / e0.elapsed_time(e1) / 1e6


for direction in ("h2d", "d2h"):
    for mb in (1, 2, 4, 8, 16, 32, 64, 128):
        a = run(mb, 1, direction)
        b = run(mb, 2, direction)
        c = run(mb, 4, direction)
        d = run(mb, 1, direction, other_dir_mb=16)
        e = run(mb, 2, direction, other_dir_mb=16)
        print(f"{direction} {mb:4d} MB: 1 stream {a:5.1f} GB/s | 2 streams {b:5.1f} | 4 streams {c:5.1f} | 1 stream + opposite 16 MB stream {d:5.1f} | 2 streams + opposite {e:5.1f}", flush=True)
