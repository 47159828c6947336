"""Runs two representative conv layers on the tensor-core path (for ncu): a full-resolution 32->32 3x3 layer over the
F*B single-frame volumes (HBM/L2-bound) and a 1/8-resolution 576->128 sub-pixel transposed conv (tensor-bound)."""
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
x = torch.randn(32, 256, 512, 32, device=dev)
if half:
    x = x.half()
conv = torch.nn.Conv2d(32, 32, 3).to(dev)
L = C.PackedConv(conv.weight, conv.bias, (32,), act=C.ACT_LEAKY, act_a=0.1)
for _ in range(3):
    y = L([x])
a, b, c = torch.randn(8, 32, 64, 192, device=dev), torch.randn(8, 32, 64, 128, device=dev), torch.randn(8, 32, 64, 256, device=dev)
if half:
    a, b, c = a.half(), b.half(), c.half()
ct = torch.nn.ConvTranspose2d(576, 128, 4, stride=2).to(dev)
R = C.refine_layer(ct, (192, 128, 256))
for _ in range(3):
    z = R([a, b, c])
torch.cuda.synchronize()
print("done", y.shape, z.shape)
