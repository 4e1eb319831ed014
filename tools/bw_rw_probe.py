"""Dev aid: write-only / read-only / copy bandwidth of this GPU with plain torch kernels (is a 4.3 GB write-only kernel bound by
the write path?  MEASURED_PEAKS.json's hbm figure is a copy: half reads, half writes)."""
import torch
n = 1 << 30                      # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
gb = n * 4 / 1e9
ms = t(lambda: a.fill_(1.0));  print(f"write-only fill_   {ms:.3f} ms  {gb/ms:.2f} TB/s")
ms = t(lambda: a.zero_());     print(f"write-only zero_   {ms:.3f} ms  {gb/ms:.2f} TB/s")
ms = t(lambda: a.sum());       print(f"read-only sum      {ms:.3f} ms  {gb/ms:.2f} TB/s")
ms = t(lambda: b.copy_(a));    print(f"copy               {ms:.3f} ms  {2*gb/ms:.2f} TB/s (read + write)")
ms = t(lambda: torch.add(a, 1.0, out=b)); print(f"add out            {ms:.3f} ms  {2*gb/ms:.2f} TB/s (read + write)")
