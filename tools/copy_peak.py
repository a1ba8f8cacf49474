"""Plain device-to-device copy bandwidth at several working-set sizes (context for the roofline
denominators: MEASURED_PEAKS.json was taken on a 4 GB working set)."""
import torch
for gb in (2, 8, 24, 48):
    n = gb * (1 << 30) // 8
    a = torch.empty(n, dtype=torch.float64, device="cuda").fill_(1.0)
    b = torch.empty_like(a)
    for _ in range(2): b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"copy of {gb} GB (working set {2*gb} GB): {ms:.2f} ms  {2*n*8/ms/1e6:.0f} GB/s (read+write)")
    del a, b
    torch.cuda.empty_cache()
