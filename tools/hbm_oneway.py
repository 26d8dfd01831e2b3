"""One-directional HBM bandwidth on this GPU with stock torch kernels (context for the tape-bound
training kernels): read-only (sum), write-only (fill) and copy over 8 GiB."""
import torch

n = 2 << 30          # 2 Gi float32 = 8 GiB
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
x.fill_(1.0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


gb = n * 4 / 1e9
print("read  (sum)   %.0f GB/s" % (gb / (timed(lambda: x.sum()) * 1e-3)))
print("write (fill)  %.0f GB/s" % (gb / (timed(lambda: y.fill_(2.0)) * 1e-3)))
print("copy  (r+w)   %.0f GB/s" % (2 * gb / (timed(lambda: y.copy_(x)) * 1e-3)))
