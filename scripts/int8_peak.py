"""Measured int8 tensor-core peak of this B200: cuBLASLt IGEMM through torch._int_mm (s8 x s8 -> s32), 8192^3.
The denominator of the Gram kernel's roofline fraction: burst = best of 10, sustained = back to back
for 3 s under the power cap.  Prints one JSON line; bench.py reads gpurun_out/int8_peak.json or re-measures (2 s).

    python scripts/int8_peak.py [--seconds 3]
"""
import argparse
import json
import time

import torch


def measure(n=8192, seconds=3.0):
    a = torch.randint(-127, 127, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-127, 127, (n, n), dtype=torch.int8, device="cuda").t().contiguous().t()  # column-major B
    ops = 2.0 * n ** 3
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, ops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            torch._int_mm(a, b)
        reps += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = ops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return {"int8_tops_burst": best, "int8_tops_sustained": sustained, "n": n, "reps_sustained": reps,
            "how": "torch._int_mm (cuBLASLt IGEMM s8 x s8 -> s32), 8192^3, best of 10 / back to back for %.0f s" % seconds}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    print(json.dumps(measure(seconds=ap.parse_args().seconds)))
