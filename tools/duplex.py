#!/usr/bin/env python
"""PCIe duplex check: H2D and D2H on separate streams at the same time, several transfer sizes."""
import torch

dev = torch.device("cuda", 0)
for mb in (4, 16, 64):
    nb = mb << 20
    h_in, h_out = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    reps = max(8, 2048 // mb)

    def run(do_in, do_out):
        a0, a1, b0, b1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        torch.cuda.synchronize()
        if do_in:
            with torch.cuda.stream(s1):
                a0.record(s1)
                for _ in range(reps):
                    d_in.copy_(h_in, non_blocking=True)
                a1.record(s1)
        if do_out:
            with torch.cuda.stream(s2):
                b0.record(s2)
                for _ in range(reps):
                    h_out.copy_(d_out, non_blocking=True)
                b1.record(s2)
        torch.cuda.synchronize()
        return (reps * nb / a0.elapsed_time(a1) / 1e6 if do_in else 0.0, reps * nb / b0.elapsed_time(b1) / 1e6 if do_out else 0.0)

    run(True, True)
    print(f"{mb} MB transfers: h2d alone {run(True, False)[0]:.1f} GB/s, d2h alone {run(False, True)[1]:.1f} GB/s, "
          f"concurrent h2d {run(True, True)[0]:.1f} + d2h {run(True, True)[1]:.1f} GB/s", flush=True)
