import torch, time
x = torch.rand(4, 8, 32, 256, 256)
xp = x.pin_memory()
print("is_pinned", xp.is_pinned(), "bytes", x.numel() * 4 / 1e6, "MB")
torch.cuda.synchronize()
for name, src in (("pinned", xp), ("pageable", x)):
    for nb in (True, False):
        ts = []
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            y = src.cuda(non_blocking=nb); torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print(name, "non_blocking", nb, "best %.1f ms -> %.1f GB/s" % (min(ts) * 1e3, x.numel() * 4 / min(ts) / 1e9))
# preallocated destination
dst = torch.empty_like(x, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    e0.record(); dst.copy_(xp, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print("copy_ into preallocated: %.1f ms" % e0.elapsed_time(e1))
