"""Shared helpers for the parity tests (test infrastructure)."""
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def u8_to_img(u8):
    return torch.from_numpy(u8.astype(np.float32)) / 255 - 0.5


def kitti_sample_dict():
    g = np.load(GOLDEN / "kitti_sample.npz")
    nF = g["frames_u8"].shape[0]
    K = torch.from_numpy(g["K"]).unsqueeze(0)
    data = {"keyframe": u8_to_img(g["keyframe_u8"]).unsqueeze(0),
            "keyframe_pose": torch.from_numpy(g["keyframe_pose"]).unsqueeze(0),
            "keyframe_intrinsics": K,
            "frames": [u8_to_img(g["frames_u8"][i]).unsqueeze(0) for i in range(nF)],
            "poses": [torch.from_numpy(g["poses"][i]).unsqueeze(0) for i in range(nF)],
            "intrinsics": [K.clone() for _ in range(nF)]}
    return data, g


def synth_small_dict(tag):
    g = np.load(GOLDEN / ("cv_synth_d64f6.npz" if tag == "d" else "cv_synth_small.npz"))
    B, nF, D, H, W, seed = [int(v) for v in g[f"{tag}_cfg"]]
    K = torch.from_numpy(g[f"{tag}_K"])
    data = {"keyframe": u8_to_img(g[f"{tag}_key_u8"]),
            "keyframe_pose": torch.eye(4).unsqueeze(0).repeat(B, 1, 1),
            "keyframe_intrinsics": K,
            "frames": [u8_to_img(g[f"{tag}_frames_u8"][i]) for i in range(nF)],
            "poses": [torch.from_numpy(g[f"{tag}_poses"][i]) for i in range(nF)],
            "intrinsics": [K.clone() for _ in range(nF)]}
    return data, D, torch.from_numpy(g[f"{tag}_cv"]), [torch.from_numpy(v) for v in g[f"{tag}_sf"]]


def compare_volumes(cv, sf, ref_cv, ref_sf, tol=1e-3, max_flip_px_per_frame=4):
    """Parity rule of SURVEY.md §8c.

    * where both agree on validity (a pixel is 'invalid' when the whole plane stack is exactly 0):
      max |delta| <= tol (north-star tolerance 1e-3 on fp32 volumes)
    * validity flips (boundary pixels whose bilinear mask sample is +-0): at most a few px per frame
    * argmax over planes identical wherever the reference's top1-top2 margin exceeds 1e-4
    Returns a dict of statistics.
    """
    stats = {}
    flips = 0
    worst = 0.0
    for f, (a, r) in enumerate(zip(sf, ref_sf)):
        za, zr = (a == 0).all(1), (r == 0).all(1)
        fl = int((za != zr).sum())
        assert fl <= max_flip_px_per_frame * a.shape[0], f"frame {f}: {fl} validity flips"
        flips += fl
        both = ~(za | zr)
        d = ((a - r).abs() * both.unsqueeze(1)).max().item()
        worst = max(worst, d)
    stats["sf_max_abs"] = worst
    stats["valid_flips"] = flips
    assert worst <= tol, f"single-frame volume max|d| = {worst}"
    za, zr = (cv == 0).all(1), (ref_cv == 0).all(1)
    both = ~(za | zr)
    stats["cv_zero_flips"] = int((za != zr).sum())
    d = ((cv - ref_cv).abs() * both.unsqueeze(1)).max().item()
    stats["cv_max_abs"] = d
    assert d <= tol, f"cost volume max|d| = {d}"
    top = torch.topk(ref_cv, 2, dim=1)[0]
    margin = top[:, 0] - top[:, 1]
    sel = both & (margin > 1e-4)
    agree_gated = (cv.argmax(1) == ref_cv.argmax(1))[sel].float().mean().item() if sel.any() else 1.0
    agree_raw = (cv.argmax(1) == ref_cv.argmax(1))[both].float().mean().item() if both.any() else 1.0
    stats["argmax_agree_gated"], stats["argmax_agree_raw"] = agree_gated, agree_raw
    assert agree_gated == 1.0, f"argmax differs on pixels with margin > 1e-4: agreement {agree_gated}"
    return stats

REPROJ_CFG = (2, 2, 48, 80, 31)    # B, F, H, W, seed of tests/golden/reprojection.npz


def reprojection_inputs():
    """Seeded inputs of the reprojection-loss golden: the synthetic dict plus a stereo frame (0.54 m baseline), a smooth
    predicted inverse depth in [0.02, 0.3] with pixel noise, and positive weights for the scalar that is differentiated.
    Shared by tests/golden/make_golden.py --only-reprojection and tests/test_reprojection.py (the inputs are re-created, not stored)."""
    B, Fn, H, W, seed = REPROJ_CFG
    from monorec_b200.synthetic import make_inputs
    d = make_inputs(B, Fn, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    d["stereoframe"] = torch.round((torch.roll(d["keyframe"], shifts=(1, 4), dims=(2, 3)) + 0.5) * 255.0) / 255.0 - 0.5
    pose = torch.eye(4).unsqueeze(0).repeat(B, 1, 1)
    pose[:, 0, 3] = 0.54
    d["stereoframe_pose"] = pose
    d["stereoframe_intrinsics"] = d["keyframe_intrinsics"].clone()
    yy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1) / H
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W) / W
    invd = 0.16 + 0.13 * torch.sin(5.0 * xx + 2.0 * yy + torch.rand(B, 1, 1, 1, generator=g) * 6.28)
    invd = (invd + 0.01 * (torch.rand(B, 1, H, W, generator=g) - 0.5)).clamp(0.02, 0.3).contiguous()
    wts = (torch.rand(B, H, W, generator=g) + 0.5).contiguous()
    return d, invd, wts
