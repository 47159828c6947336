"""Times the full MonoRecModel forward and its stages on cuda:0 (CUDA events), synthetic KITTI-shaped inputs."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from monorec_b200.model import MonoRecModel  # noqa: E402
from monorec_b200.synthetic import make_inputs, to_device  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
F = int(sys.argv[2]) if len(sys.argv) > 2 else 4
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(0)
model = MonoRecModel().cuda().eval()
data = to_device(make_inputs(B, F, 256, 512, seed=0), "cuda:0")


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    d = dict(data)
    t_all = timed(lambda: model(dict(data)), iters)
    d = model(dict(data))
    d["_cv_range"] = (0.0025, 0.33, 32)
    t_cv = timed(lambda: model.cv_module(d), iters)
    t_res = timed(lambda: model._feature_extractor(d["keyframe"] + .5), iters)
    t_mask = timed(lambda: model.att_module(d), iters)
    d["_cv_mask_for_depth"] = d["cv_mask"]
    t_depth = timed(lambda: model.depth_module(d), iters)
print(f"B={B} F={F}: forward {t_all:.2f} ms ({1e3 * B / t_all:.1f} keyframes/s) | cost volume {t_cv:.2f} | resnet {t_res:.2f} | "
      f"mask {t_mask:.2f} | depth {t_depth:.2f} ms")
