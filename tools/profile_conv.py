"""Runs representative conv layers on the tensor-core path (for ncu): a full-resolution 32->32 3x3 layer over the F*B
single-frame volumes (resident-weight halo kernel), the 32+64->48 3x3 decoder layer at full resolution (halo kernel with streamed
weights), a 1/8-resolution 576->128 and a 1/2-resolution 192->48 sub-pixel transposed conv (tap-refetch kernel, four phases in
one launch)."""
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
p, q = torch.randn(8, 256, 512, 32, device=dev), torch.randn(8, 256, 512, 64, device=dev)
a2 = [torch.randn(8, 128, 256, 64, device=dev) for _ in range(3)]
if half:
    p, q, a2 = p.half(), q.half(), [t.half() for t in a2]
conv2 = torch.nn.Conv2d(96, 48, 3).to(dev)
S = C.PackedConv(conv2.weight, conv2.bias, (32, 64), act=C.ACT_LEAKY, act_a=0.1)
R2 = C.refine_layer(torch.nn.ConvTranspose2d(192, 48, 4, stride=2).to(dev), (64, 64, 64))
for _ in range(3):
    w = S([p, q])
    v = R2(a2)
torch.cuda.synchronize()
print("done", y.shape, z.shape, w.shape, v.shape)
