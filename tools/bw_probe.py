"""Developer probe: pure-write, pure-read and copy HBM bandwidth on this GPU with torch (fill_, sum, copy_), 8 GiB buffers."""
import torch, time
n = 1 << 30
a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty(n, dtype=torch.float64, device="cuda")
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("write  %.0f GB/s" % (n * 8 / t(lambda: a.fill_(1.0)) / 1e6))
print("read   %.0f GB/s" % (n * 8 / t(lambda: a.sum()) / 1e6))
print("copy   %.0f GB/s (read+write)" % (2 * n * 8 / t(lambda: b.copy_(a)) / 1e6))
