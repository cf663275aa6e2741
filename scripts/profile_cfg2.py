"""Kernel-time breakdown of the cfg-2 training step (8 images x 2048 queries, fp32, forward + backward) with torch's
profiler: which kernels the 12 ms go to (device time, summed over the step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench_configs as BC  # noqa: E402


def main():
    step = BC.cfg2(return_step=True)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    total = sum(e.device_time_total for e in rows)
    print(f"device time per step: {total / 3e3:.3f} ms")
    for e in rows[:25]:
        print(f"{e.device_time_total / 3e3:8.3f} ms  x{e.count // 3:4d}  {e.key[:110]}")


if __name__ == "__main__":
    main()
