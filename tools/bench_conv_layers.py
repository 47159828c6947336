"""Times representative conv layers of the stacks on the tensor-core path (CUDA events), halo variant on/off via env."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from monorec_b200 import conv as C  # noqa: E402

import os  # noqa: E402

MODE = os.environ.get("MONOREC_B200_CONV", "tf32")
C.set_mode(MODE)
half = MODE == "f16"
dev = "cuda:0"
torch.manual_seed(0)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


cases = [("mask enc0 3x3 32->32 B32 full", 32, 256, 512, (32,), 32, 3, 3),
         ("depth enc0 7x1 40->48 B8 full", 8, 256, 512, (40,), 48, 7, 1),
         ("depth enc0 1x7 48->48 B8 full", 8, 256, 512, (48,), 48, 1, 7),
         ("depth enc0 3x1 48->48 B8 full", 8, 256, 512, (48,), 48, 3, 1),
         ("depth enc0 1x3 48->48 B8 full", 8, 256, 512, (48,), 48, 1, 3),
         ("depth dec4 3x1 48+48->32 B8 full", 8, 256, 512, (48, 48), 32, 3, 1),
         ("depth dec4 1x3 32->32 B8 full", 8, 256, 512, (32,), 32, 1, 3),
         ("depth dec4 3x3 32->24 B8 full", 8, 256, 512, (32,), 24, 3, 3),
         ("depth enc1 3x1 64->64 B8 half", 8, 128, 256, (64,), 64, 3, 1),
         ("mask dec3.2 3x3 48->48 B8 full", 8, 256, 512, (48,), 48, 3, 3),
         ("mask dec3.1 3x3 32+64->48 B8 full", 8, 256, 512, (32, 64), 48, 3, 3),
         ("mask enc1 3x3 48->48 B32 half", 32, 128, 256, (48,), 48, 3, 3),
         ("mask dec1.1 3x3 64+64+96->96 B8 1/4", 8, 64, 128, (64, 64, 96), 96, 3, 3)]
for name, B, H, W, src_c, cout, kh, kw in cases:
    xs = [torch.randn(B, H, W, c, device=dev) for c in src_c]
    if half:
        xs = [x.half() for x in xs]
    conv = torch.nn.Conv2d(sum(src_c), cout, (kh, kw)).to(dev)
    L = C.PackedConv(conv.weight, conv.bias, src_c, act=C.ACT_LEAKY, act_a=0.1)
    us = timed(lambda: L(xs))
    flops = 2.0 * B * H * W * sum(src_c) * cout * kh * kw
    byts = (2.0 if half else 4.0) * B * H * W * (sum(src_c) + cout)
    print(f"{name:40s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {byts / us / 1e3:7.1f} GB/s (in+out)")

# sub-pixel layers (four phases): Refine = ConvTranspose2d(k4, s2) + crop, Upconv = nearest x2 + 2x2 conv
for name, B, H, W, src_c, cout, kind in [("refine 64+64+64->48 B8 1/2", 8, 128, 256, (64, 64, 64), 48, "refine"),
                                         ("refine 128+64+128->64 B8 1/4", 8, 64, 128, (128, 64, 128), 64, "refine"),
                                         ("refine 192+128+256->128 B8 1/8", 8, 32, 64, (192, 128, 256), 128, "refine"),
                                         ("refine 256->256 B8 1/16", 8, 16, 32, (256,), 256, "refine"),
                                         ("upconv 64->64 B8 1/2", 8, 128, 256, (64,), 64, "upconv"),
                                         ("upconv 96->96 B8 1/4", 8, 64, 128, (96,), 96, "upconv")]:
    xs = [torch.randn(B, H, W, c, device=dev) for c in src_c]
    if half:
        xs = [x.half() for x in xs]
    if kind == "refine":
        L = C.refine_layer(torch.nn.ConvTranspose2d(sum(src_c), cout, 4, 2).to(dev), src_c)
    else:
        L = C.upconv_layer(torch.nn.Conv2d(sum(src_c), cout, 2).to(dev), src_c)
    us = timed(lambda: L(xs))
    es = 2.0 if half else 4.0
    byts = es * B * H * W * (sum(src_c) + 4 * cout)
    print(f"{name:40s} {us:8.1f} us  {byts / us / 1e3:7.1f} GB/s (in once + out)")
