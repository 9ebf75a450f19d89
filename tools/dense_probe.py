#!/usr/bin/env python
"""Timing probe of the dense path (config 5): Gram kernel and Newton-loop kernel times for one batch."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=4000)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    import torch
    from tenscalc_b200.dense import DenseLSSolver, gram_flops, newton_step_flops
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((a.batch, a.n, a.N), generator=g, device="cuda", dtype=torch.float64) / np.sqrt(a.N)   # [B][col][row] = column-major
    b = torch.randn((a.batch, a.N), generator=g, device="cuda", dtype=torch.float64)
    x0 = torch.zeros((a.batch, a.n), device="cuda", dtype=torch.float64)
    torch.cuda.synchronize()
    obj = DenseLSSolver(a.N, a.n, a.batch, -1.0, 1.0)
    for r in range(a.repeat):
        obj.set_device_ptrs(A.data_ptr(), b.data_ptr(), x0.data_ptr())
        t = time.time()
        st, it, sec = obj.solve(1.0, 50, -1)
        gm, sm = obj.times_ms()
        steps = int(it.sum())
        print(f"B={a.batch} N={a.N} n={a.n}: gram {gm:.2f} ms ({gram_flops(a.N, a.n) * a.batch / gm / 1e9:.2f} TFLOP/s algorithmic), "
              f"newton loop {sm:.2f} ms ({steps} steps, {sm * 1e3 / max(steps, 1):.2f} us/step, "
              f"{newton_step_flops(a.n) * steps / sm / 1e9:.2f} TFLOP/s algorithmic), wall {1e3 * (time.time() - t):.1f} ms, "
              f"status {dict(zip(*np.unique(st, return_counts=True)))}, iters {it.min()}..{it.max()}, "
              f"{a.batch / (gm + sm) * 1e3:.0f} solves/s")
    print("counters", obj.counters())
    obj.close()


if __name__ == "__main__":
    main()
