#!/bin/bash
# compute-sanitizer memcheck on the conv engine's three kernel shapes of this round (resident halo with 8..14-px rows, streamed
# weights, sub-pixel phases in one launch, one-channel head epilogue) and on the reprojection loss
mkdir -p gpurun_out
cat > /tmp/san_k2.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from monorec_b200 import conv as C
from monorec_b200 import losses as L
from monorec_b200.synthetic import make_inputs, to_device
torch.manual_seed(0)
dev = "cuda:0"
for mode in ("tf32", "f16"):
    C.set_mode(mode)
    half = mode == "f16"
    cast = (lambda t: t.half()) if half else (lambda t: t)
    for (cin, cout, kh, kw, H, W) in [((32,), 32, 3, 3, 21, 37), ((40,), 48, 7, 1, 24, 40), ((48,), 48, 1, 7, 24, 40), ((32, 64), 48, 3, 3, 20, 36),
                                      ((48, 64, 96), 64, 3, 3, 18, 20), ((96,), 96, 3, 3, 9, 13), ((128,), 128, 3, 1, 12, 16), ((24,), 1, 3, 3, 17, 33)]:
        xs = [cast(torch.randn(2, H, W, c, device=dev)) for c in cin]
        conv = torch.nn.Conv2d(sum(cin), cout, (kh, kw)).to(dev)
        act = C.ACT_ABSTANH if cout == 1 else C.ACT_LEAKY
        y = C.PackedConv(conv.weight, conv.bias, cin, act=act, act_a=0.1, act_b=1.0)(xs, final=(cout == 1))
    x = cast(torch.randn(2, 12, 20, 64, device=dev))
    z = C.refine_layer(torch.nn.ConvTranspose2d(64, 48, 4, stride=2).to(dev), (64,))([x])
    u = C.upconv_layer(torch.nn.Conv2d(64, 64, 2).to(dev), (64,))([x])
    torch.cuda.synchronize()
d = to_device(make_inputs(1, 3, 50, 70, seed=5), dev)
invd = (0.1 + 0.1 * torch.rand(1, 1, 50, 70, device=dev)).requires_grad_(True)
L.reprojection_loss(invd, d, automasking=True, reduce=True).backward()
L.reprojection_loss(invd, d, border=2, reduce=False)
torch.cuda.synchronize()
print("k2 + loss done", tuple(z.shape), tuple(u.shape), float(invd.grad.abs().max()))
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_k2.py > gpurun_out/san_k2_mem.log 2>&1; echo "k2 memcheck exit $?"; tail -4 gpurun_out/san_k2_mem.log | cut -c1-200
