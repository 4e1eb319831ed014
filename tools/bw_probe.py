"""Dev aid: read / write / copy HBM bandwidth of this GPU with plain torch kernels (CUDA events, best of 10)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
b = torch.empty(n, dtype=torch.float32, device="cuda")
def best(f, reps=10):
    f(); torch.cuda.synchronize(); t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); t.append(e0.elapsed_time(e1))
    return min(t)
gb = n * 4 / 1e9
print("write  (fill_)      %.0f GB/s" % (gb / best(lambda: b.fill_(1.0)) * 1e3))
print("write  (zero_)      %.0f GB/s" % (gb / best(lambda: b.zero_()) * 1e3))
print("read   (sum)        %.0f GB/s" % (gb / best(lambda: a.sum()) * 1e3))
print("copy   (copy_)      %.0f GB/s  (read + write bytes)" % (2 * gb / best(lambda: b.copy_(a)) * 1e3))
print("r2w1   (a+b->b)     %.0f GB/s" % (3 * gb / best(lambda: torch.add(a, b, out=b)) * 1e3))
print("mul    (b*=2)       %.0f GB/s  (in place: read + write same lines)" % (2 * gb / best(lambda: b.mul_(2.0)) * 1e3))
