import time, pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
import torch; torch.zeros(1, device="cuda"); torch.cuda.synchronize()
for name, f in (("clock", lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                ("reasons", lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h))):
    ts = []
    for _ in range(20):
        t = time.perf_counter(); f(); ts.append((time.perf_counter() - t) * 1e3)
    print(name, "ms min/med/max", min(ts), sorted(ts)[10], max(ts))
