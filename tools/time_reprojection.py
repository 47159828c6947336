#!/usr/bin/env python
"""CUDA-event timing of the reprojection loss (forward, forward + backward) at BASELINE config 2's shapes, next to the same
loss evaluated by stock PyTorch CUDA ops written like the reference (the oracle's primitives on the device)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from monorec_b200 import losses as L  # noqa: E402
from monorec_b200.synthetic import make_inputs, to_device  # noqa: E402


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B, Fn, H, W = 8, 4, 256, 512
    d = to_device(make_inputs(B, Fn, H, W, seed=0), "cuda:0")
    invd = (0.15 + 0.1 * torch.rand(B, 1, H, W, device="cuda:0")).requires_grad_(True)

    def fwd():
        with torch.no_grad():
            return L.reprojection_loss(invd, d, automasking=True, reduce=False)

    def fwd_bwd():
        invd.grad = None
        L.reprojection_loss(invd, d, automasking=True, reduce=True).backward()

    t1, t2 = timeit(fwd), timeit(fwd_bwd)
    px = B * H * W
    print(f"reprojection_loss B={B} F={Fn} {H}x{W} automasking: forward {t1:.3f} ms, forward+backward {t2:.3f} ms "
          f"({px / t2 / 1e3:.0f} Mpx/s)")


if __name__ == "__main__":
    main()
