"""Measured FP64 peaks of this B200 (SURVEY 8d asks for a DGEMM figure: the roofline of the dense supernode updates).
  dgemm: cuBLAS through torch.matmul, 8192^3 doubles (2 N^3 flop): best of 10 (burst) and back to back for 3 s (sustained)
Prints one JSON object (committed as profiles/fp64_peak.json)."""
import json
import time

import torch


def main():
    dev = torch.device("cuda", 0)
    out = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        c = torch.empty_like(a)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["dgemm_%d_burst_tflops" % n] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 0
        e0.record()
        while time.time() - t0 < 3.0:
            for _ in range(4):
                torch.matmul(a, b, out=c)
            reps += 4
            torch.cuda.synchronize()
        e1.record()
        e1.synchronize()
        out["dgemm_%d_sustained_tflops" % n] = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    out["how"] = "torch.matmul float64 (cuBLAS DGEMM), CUDA events"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
