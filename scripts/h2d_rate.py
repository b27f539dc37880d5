"""Host -> device copy rate of this box from page-locked and from pageable memory (torch, CUDA events), for DESIGN.md 7.2."""
import torch

for mb in (4, 16, 64):
    n = mb << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, host in (("page-locked", torch.empty(n, dtype=torch.uint8).pin_memory()), ("pageable", torch.empty(n, dtype=torch.uint8))):
        host.fill_(1)
        best = 1e9
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            dev.copy_(host, non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print("%3d MiB %-11s %.3f ms  %.1f GB/s" % (mb, name, best, n / best / 1e6))

# the solver's case: the page-locked source was written by many host threads just before the copy (dirty lines in many caches)
import numpy as np  # noqa: E402

n = 16 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
for threads in (1, 16, 64):
    torch.set_num_threads(threads)
    times = []
    for it in range(6):
        host.add_(1)  # rewritten by `threads` host threads (intra-op parallelism)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        dev.copy_(host, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    print("16 MiB page-locked, rewritten by %2d threads before every copy: median %.3f ms  %.1f GB/s" % (threads, float(np.median(times)), n / float(np.median(times)) / 1e6))
