"""CUDA-event timing of the cost-volume kernel alone (events bracket each mr_cost_volume_fwd call; outputs preallocated).

    python tools/time_cv.py [B F D H W iters]      (MONOREC_B200_LIB selects a variant build)
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from monorec_b200 import _lib  # noqa: E402
from monorec_b200.synthetic import make_inputs, to_device  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
B, F, D, H, W, iters = (a + [8, 4, 32, 256, 512, 10][len(a):])[:6]
dev = "cuda:0"
lib = _lib.load()
sets = [to_device(make_inputs(B, F, H, W, seed=s), dev) for s in range(3)]
proj = torch.empty(B, F, 3, 4, device=dev)
depths = torch.empty(D, device=dev)
cv = torch.empty(B, D, H, W, device=dev)
sfcv = torch.empty(F, B, D, H, W, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for name in ("mr_cost_volume_fwd", "mr_cost_volume_fwd_gather"):
    fn = getattr(lib, name)
    times = []
    for i in range(iters + 3):
        d = sets[i % 3]
        _lib.check(lib.mr_projection_tables(d["keyframe_pose"].data_ptr(), d["keyframe_intrinsics"].data_ptr(),
                                            _lib.ptr_array(d["poses"]), _lib.ptr_array(d["intrinsics"]), B, F, H, W,
                                            proj.data_ptr(), depths.data_ptr(), D, 0.0025, 0.33, stream), "tables")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(fn(d["keyframe"].data_ptr(), _lib.ptr_array(d["frames"]), proj.data_ptr(), depths.data_ptr(),
                      cv.data_ptr(), sfcv.data_ptr(), B, F, D, H, W, 10.0, None, stream), name)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    alg = 4.0 * H * W * (1 + F) * (3 + D) * B
    print(f"B={B} F={F} D={D} {H}x{W} {name}: {ms:.3f} ms (min {min(times):.3f}), {alg / ms / 1e6:.1f} GB/s algorithmic")
