"""Diagnostics: does in-process NVML work on the box, and what does one poll cost?"""
import time, traceback
import torch
try:
    import pynvml
    pynvml.nvmlInit()
    print("count", pynvml.nvmlDeviceGetCount())
    p = torch.cuda.get_device_properties(0)
    print("torch pci_bus_id", getattr(p, "pci_bus_id", None))
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    print("pci", pynvml.nvmlDeviceGetPciInfo(h).bus)
    for _ in range(3):
        t0 = time.perf_counter()
        a = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        b = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        c = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        print("poll", a, b, hex(c), f"{(time.perf_counter() - t0) * 1e3:.2f} ms")
except Exception:
    traceback.print_exc()
