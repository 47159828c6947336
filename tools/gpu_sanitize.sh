#!/bin/bash
# compute-sanitizer memcheck / racecheck on small configurations of the two kernel families
mkdir -p gpurun_out
cat > /tmp/san_k1.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from monorec_b200.cost_volume import CostVolumeModule
from monorec_b200.synthetic import make_inputs, to_device
for (B, F, D, H, W) in [(1, 2, 8, 37, 61), (1, 3, 32, 48, 332), (2, 2, 16, 40, 132), (1, 2, 32, 64, 128)]:
    d = to_device(make_inputs(B, F, H, W, seed=3), "cuda:0"); d["_cv_range"] = (0.0025, 0.33, D)
    for tma in (True, False):
        m = CostVolumeModule(); m.tma_windows = tma
        o = m(d); torch.cuda.synchronize()
print("k1 done")
PY
cat > /tmp/san_k2.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from monorec_b200 import conv as C
torch.manual_seed(0)
for mode in ("fp32", "tf32"):
    C.set_mode(mode)
    x = torch.randn(2, 24, 40, 36, device="cuda:0")
    conv = torch.nn.Conv2d(36, 48, (7, 1), stride=(2, 1)).cuda()
    y = C.PackedConv(conv.weight, conv.bias, (36,), stride=(2, 1), act=C.ACT_LEAKY, act_a=0.1)([x])
    ct = torch.nn.ConvTranspose2d(48, 24, 4, stride=2).cuda()
    z = C.refine_layer(ct, (48,))([y])
    h = torch.nn.Conv2d(24, 1, 3).cuda()
    w = C.PackedConv(h.weight, h.bias, (24,), act=C.ACT_ABSTANH, act_a=0.0, act_b=1.0, allow_tc=False)([z], final=True)
    torch.cuda.synchronize()
print("k2 done", w.shape)
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_k1.py > gpurun_out/san_k1_mem.log 2>&1; echo "k1 memcheck exit $?"; tail -3 gpurun_out/san_k1_mem.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_k1.py > gpurun_out/san_k1_race.log 2>&1; echo "k1 racecheck exit $?"; tail -3 gpurun_out/san_k1_race.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_k2.py > gpurun_out/san_k2_mem.log 2>&1; echo "k2 memcheck exit $?"; tail -3 gpurun_out/san_k2_mem.log
