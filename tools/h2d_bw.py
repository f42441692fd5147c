import torch, time
n = 1 << 30
h = torch.empty(n // 4, dtype=torch.float32, pin_memory=True); d = torch.empty(n // 4, dtype=torch.float32, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    print(name, "GB/s", 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
