import torch, time
for mb in (1, 2, 4, 16):
    n = mb * (1 << 20) // 8
    h = torch.empty(n, dtype=torch.double).pin_memory(); d = torch.empty(n, dtype=torch.double, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {mb} MiB pinned: {ms*1e3:.1f} us, {mb*1.048576/ms:.1f} GB/s")
