"""Calibration: achievable HBM bandwidth of write-only / copy / read-only streams on this GPU (torch library kernels).
Used to put the obs-write-dominated step kernel (83 % of its traffic is writes) in context; the roofline
denominator stays MEASURED_PEAKS.json's copy figure."""
import json
import torch

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3

def main():
    n = 5 * 400 * 1024 ** 2      # 2000 Mi floats = 7.8 GiB
    a = torch.empty(n, dtype=torch.float32, device="cuda")
    b = torch.empty(n, dtype=torch.float32, device="cuda")
    a.fill_(1.0)
    res = {}
    res["write_only_fill_GBs"] = n * 4 / timeit(lambda: b.fill_(2.0)) / 1e9
    res["write_only_zero_GBs"] = n * 4 / timeit(lambda: b.zero_()) / 1e9
    res["copy_rw_GBs"] = 2 * n * 4 / timeit(lambda: b.copy_(a)) / 1e9
    res["read_only_sum_GBs"] = n * 4 / timeit(lambda: a.sum()) / 1e9
    # 1 read : 5 writes, like the step kernel (ring read 200 B + obs write 1000 B per asset-step)
    c = a[: n // 5]
    res["mix_1r5w_GBs"] = (n // 5 * 4 + n // 5 * 5 * 4) / timeit(lambda: b.view(5, -1)[:, : n // 5].copy_(c.expand(5, -1))) / 1e9
    print(json.dumps(res))

if __name__ == "__main__":
    main()
