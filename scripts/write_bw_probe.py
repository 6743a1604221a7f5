import torch
x = torch.empty(1_000_000, 1096, device="cuda")
for name, fn in (("zero_", lambda: x.zero_()), ("fill_", lambda: x.fill_(1.5)), ("copy_ (r+w)", None)):
    if fn is None:
        y = torch.empty_like(x)
        fn = lambda: y.copy_(x)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = x.numel() * 4 / 1e9 * (2 if "copy" in name else 1)
    print(name, round(ms, 3), "ms", round(gb / ms * 1e3), "GB/s")
