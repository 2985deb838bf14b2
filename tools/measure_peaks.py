#!/usr/bin/env python
"""One-off yardsticks the driver's MEASURED_PEAKS.json does not hold: cuBLAS TF32 and FP32 (FFMA) 8192^3 throughput on
this pool's B200, through torch.matmul.  NEVER on the product path — bench.py only quotes the numbers next to its own
(roofline.cublas_yardstick) and uses the FFMA figure as the denominator of the exact-fp32 mode.

    python tools/measure_peaks.py > profiles/r2_peaks.json
"""
import json
import time

import torch


def rate(dtype_cfg, n=8192, burst_reps=10, sustain_s=3.0):
    torch.backends.cuda.matmul.allow_tf32 = dtype_cfg == "tf32"
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(burst_reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < sustain_s:
        for _ in range(8):
            a @ b
        reps += 8
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return round(best, 1), round(2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12, 1)


def main():
    tf32_burst, tf32_sus = rate("tf32")
    fp32_burst, fp32_sus = rate("fp32", burst_reps=5, sustain_s=2.0)
    print(json.dumps({"how": "torch.matmul fp32 8192^3 through cuBLAS: allow_tf32 on (TF32 tensor cores) / off (FFMA); best of N (burst) and back to back for seconds (sustained)",
                      "gpu": torch.cuda.get_device_name(0), "tf32_tflops": tf32_burst, "tf32_tflops_sustained": tf32_sus,
                      "fp32_ffma_tflops": fp32_burst, "fp32_ffma_tflops_sustained": fp32_sus}, indent=1))


if __name__ == "__main__":
    main()
