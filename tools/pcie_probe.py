import torch, time
dev = torch.device("cuda", 0)
for mb in (1, 5, 20, 64, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(3)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print(f"{mb:4d} MiB  H2D {ms:.3f} ms {n/ms/1e6:.1f} GB/s   D2H {ms2:.3f} ms {n/ms2/1e6:.1f} GB/s")
