#!/usr/bin/env python
"""Writes the CUDA reprojection-loss outputs (errors, winners, gradients) of the test configurations to gpurun_out/reproj_dump.npz
so that outliers can be analysed on the dev container against the float64 closed form."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests.test_reprojection import CASES, _cuda_run  # noqa: E402
from tests.helpers import reprojection_inputs  # noqa: E402
from monorec_b200.synthetic import make_inputs  # noqa: E402

out = {}
data, invd, wts = reprojection_inputs()
for tag, kw in CASES.items():
    e, win, grad = _cuda_run(invd, data, wts, **kw)
    out[f"e_{tag}"], out[f"w_{tag}"], out[f"g_{tag}"] = e.numpy(), win.numpy(), grad.numpy()
for i, cfg in enumerate([(1, 3, 50, 70, 5, dict(automasking=True)), (1, 4, 256, 512, 100, dict(automasking=True))]):
    B, Fn, H, W, seed, kw = cfg
    d = make_inputs(B, Fn, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed)
    yy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1) / H
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W) / W
    iv = (0.15 + 0.12 * torch.sin(4.0 * xx + 3.0 * yy) + 0.01 * (torch.rand(B, 1, H, W, generator=g) - 0.5)).clamp(0.01, 0.3)
    wt = torch.rand(B, H, W, generator=g) + 0.5
    e, win, grad = _cuda_run(iv, d, wt, **kw)
    out[f"e_cfg{i}"], out[f"w_cfg{i}"], out[f"g_cfg{i}"] = e.numpy(), win.numpy(), grad.numpy()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
np.savez_compressed(ROOT / "gpurun_out" / "reproj_dump.npz", **out)
print("dumped", sorted(out))
