#!/bin/bash
# compute-sanitizer memcheck over whole MonoRecModel forwards (every layer configuration of the real network, both
# tensor-core arithmetic modes, a small and the full image size, graph replay excluded)
mkdir -p gpurun_out
cat > /tmp/san_model.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from monorec_b200 import conv as C
from monorec_b200.model import MonoRecModel
from monorec_b200.synthetic import make_inputs, to_device
torch.manual_seed(0)
model = MonoRecModel().cuda().eval()
for mode in ("f16", "tf32"):
    C.set_mode(mode)
    for (B, F, H, W) in [(1, 2, 64, 128), (2, 3, 48, 80), (1, 4, 256, 512)]:
        with torch.no_grad():
            out = model(to_device(make_inputs(B, F, H, W, seed=1), "cuda:0"))
        torch.cuda.synchronize()
        print(mode, (B, F, H, W), float(out["result"].float().mean()), len(out["image_features"]))
print("model done")
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_model.py > gpurun_out/san_model_mem.log 2>&1; echo "model memcheck exit $?"; tail -9 gpurun_out/san_model_mem.log | cut -c1-160
