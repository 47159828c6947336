"""Runs the fused cost-volume kernel a few times at BASELINE config 2 (for `ncu -k regex:cost_volume`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from monorec_b200.cost_volume import CostVolumeModule  # noqa: E402
from monorec_b200.synthetic import make_inputs, to_device  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = int(sys.argv[2]) if len(sys.argv) > 2 else 4
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
d = to_device(make_inputs(B, F, 256, 512, seed=0), "cuda:0")
d["_cv_range"] = (0.0025, 0.33, 32)
m = CostVolumeModule()
for _ in range(iters):
    m(d)
torch.cuda.synchronize()
print("done")
