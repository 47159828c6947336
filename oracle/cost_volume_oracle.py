"""ORACLE (test infrastructure, not product code) -- CPU restatement of MonoRec's plane-sweep cost volume.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file, and only as the checker / the CPU baseline.  The product path (monorec_b200/) never
imports anything from `oracle/`.

Reference being restated: /root/reference/model/monorec/monorec_model.py:150-284 (CostVolumeModule.forward,
create_mask) with model/layers.py:43-71 (Backprojection, point_projection) and :91-139 (SSIM).

Parity pin: the reference has no tests or golden vectors of its own ("parity unpinned" by the reference,
SURVEY.md §4/§8c).  This oracle is pinned instead against outputs of the *reference itself* run in the dev
container (tests/golden/make_golden.py imports it unmodified from /root/reference and writes
tests/golden/*.npz); tests/test_oracle_golden.py checks both restatements below against those files.

Two independent restatements:

* `cost_volume_torch`  -- same library primitives as the reference (F.grid_sample, avg_pool2d, conv3d), so it
  has the reference's CPU performance characteristics; this is what `bench.py` times as the CPU baseline
  ("port").
* `cost_volume_closed_form` -- numpy, explicit bilinear gather and box sums following SURVEY.md Appendix C;
  shares no primitive with the first one and can run in float64 (used for tie margins).
"""
import numpy as np
import torch
import torch.nn.functional as F

CHANNEL_WEIGHTS = (5.0 / 32.0, 16.0 / 32.0, 11.0 / 32.0)  # monorec_model.py:133
ALPHA = 10.0                                                # monorec_model.py:133
SSIM_C1 = 0.01 ** 2                                         # layers.py:116
SSIM_C2 = 0.03 ** 2                                         # layers.py:117


def plane_depths(inv_depth_min, inv_depth_max, steps, dtype=torch.float32):
    """z_d = 1 / linspace(inv_depth_max_value, inv_depth_min_value, D)  (monorec_model.py:184-185).

    Note the reference's names are swapped w.r.t. their values: data_dict["inv_depth_max"] holds the
    *smaller* number (0.0025), so index 0 is the farthest plane (400 m).
    """
    return 1.0 / torch.linspace(float(inv_depth_max), float(inv_depth_min), int(steps), dtype=dtype)


def collect_frames(data, use_mono=True, use_stereo=False):
    """monorec_model.py:156-167."""
    frames, intrinsics, poses = [], [], []
    if use_mono:
        frames += list(data["frames"])
        intrinsics += list(data["intrinsics"])
        poses += list(data["poses"])
    if use_stereo:
        frames.append(data["stereoframe"])
        intrinsics.append(data["stereoframe_intrinsics"])
        poses.append(data["stereoframe_pose"])
    return frames, intrinsics, poses


def interior_mask(height, width, border, dtype=torch.float32):
    """1 inside, 0 in a `border`-pixel ring (monorec_model.py:282-284)."""
    m = torch.zeros(1, 1, height, width, dtype=dtype)
    m[:, :, border:height - border, border:width - border] = 1
    return m


def _pixel_grid(height, width, dtype):
    # layers.py:49-54: rows of [x; y; 1], row-major over (y, x)
    ys, xs = torch.meshgrid(torch.arange(height, dtype=dtype), torch.arange(width, dtype=dtype), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, dtype=dtype)], 0)


def _ssim_error(x, y):
    """layers.py:119-137 with the default ctor (reflection pad 1, 3x3 average pools)."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    mu_xx, mu_yy, mu_xy = mu_x ** 2, mu_y ** 2, mu_x * mu_y
    sig_x = F.avg_pool2d(x ** 2, 3, 1) - mu_xx
    sig_y = F.avg_pool2d(y ** 2, 3, 1) - mu_yy
    sig_xy = F.avg_pool2d(x * y, 3, 1) - mu_xy
    num = (2 * mu_xy + SSIM_C1) * (2 * sig_xy + SSIM_C2)
    den = (mu_xx + mu_yy + SSIM_C1) * (sig_x + sig_y + SSIM_C2)
    return torch.clamp((1 - num / den) / 2, 0, 1)


@torch.no_grad()
def cost_volume_torch(data, inv_depth_min=0.33, inv_depth_max=0.0025, steps=32, use_mono=True, use_stereo=False,
                      patch_size=3, alpha=ALPHA, channel_weights=CHANNEL_WEIGHTS, return_valid=False):
    """Restates CostVolumeModule.forward (use_ssim=True, sfcv_mult_mask=True, not_center_cv=False).

    Returns (cost_volume (B,D,H,W), [F x (B,D,H,W)] single-frame volumes[, valid (B,F,H,W)]).
    """
    key = data["keyframe"]
    dtype = key.dtype
    frames, intrinsics, poses = collect_frames(data, use_mono, use_stereo)
    B, C, H, W = key.shape
    nF = len(frames)
    D = int(steps)
    border = patch_size // 2 + 1                                        # monorec_model.py:139
    z = plane_depths(inv_depth_min, inv_depth_max, D, dtype)            # (D,)
    grid_px = _pixel_grid(H, W, dtype)                                  # (3, HW)
    inside = interior_mask(H, W, border, dtype)
    sad_w = (torch.tensor(channel_weights, dtype=dtype) / patch_size ** 2).view(1, C, 1, 1, 1) \
        .repeat(1, 1, 1, patch_size, patch_size)                        # monorec_model.py:141-142

    out_cv, out_sf, out_valid = [], [[] for _ in range(nF)], []
    for b in range(B):                                                  # monorec_model.py:193
        kinv = torch.inverse(data["keyframe_intrinsics"][b])[:3, :3]
        rays = kinv @ grid_px                                           # (3, HW)
        pts = z.view(D, 1, 1) * rays.unsqueeze(0)                       # (D, 3, HW)   :199-200
        pts = torch.cat([pts, torch.ones(D, 1, H * W, dtype=dtype)], 1)  # homogeneous :201
        warped, valid = [], []
        for f in range(nF):
            T = torch.inverse(poses[f][b]) @ data["keyframe_pose"][b]   # :171,207
            P = (intrinsics[f][b] @ T)[:3, :]                           # layers.py:65
            cam = P.unsqueeze(0) @ pts                                  # (D, 3, HW)
            uv = cam[:, :2] / (cam[:, 2:3] + 1e-7)                      # layers.py:66
            uv = torch.stack([uv[:, 0] / (W - 1), uv[:, 1] / (H - 1)], 1)
            g = ((uv - 0.5) * 2).view(D, 2, H, W).permute(0, 2, 3, 1).clamp(-2, 2)   # :67-70, monorec :208
            img = frames[f][b:b + 1].expand(D, -1, -1, -1)
            warped.append(F.grid_sample(img, g, mode="bilinear", padding_mode="zeros", align_corners=False))
            hit = F.grid_sample(inside.expand(D, -1, -1, -1), g, mode="bilinear", padding_mode="zeros",
                                align_corners=False)
            valid.append(inside[0] * torch.min(hit != 0, dim=0)[0])     # :218-219  (1,H,W)
        warped = torch.stack(warped, 1)                                 # (D, F, C, H, W)
        valid = torch.stack(valid)                                      # (F, 1, H, W)
        n = D * nF
        err = _ssim_error(warped.reshape(n, C, H, W) + 0.5, key[b:b + 1].expand(n, -1, -1, -1) + 0.5)
        err = err.view(D, nF, C, H, W).permute(1, 2, 0, 3, 4)           # (F, C, D, H, W)
        sad = F.conv3d(err, sad_w, padding=(0, patch_size // 2, patch_size // 2)).squeeze(1)  # (F, D, H, W)
        sfcv = (1 - sad * 2) * valid                                    # :251
        for f in range(nF):
            out_sf[f].append(sfcv[f])
        spread = torch.exp(-alpha * (sad - sad.min(dim=1, keepdim=True)[0]) ** 2)   # :257
        wgt = 1 - (spread.sum(dim=1, keepdim=True) - 1) / (D - 1)       # :258
        wgt = wgt * valid                                               # :260
        num = (sad * wgt).sum(0)                                        # (D, H, W)
        den = wgt.sum(0).squeeze(0)                                     # (H, W)
        nz = den != 0
        cv = torch.zeros_like(num)
        cv[:, nz] = 1 - 2 * (num[:, nz] / den[nz])                      # :266-269
        out_cv.append(cv)
        out_valid.append(valid[:, 0])
    cost_volume = torch.stack(out_cv)
    single = [torch.stack(v) for v in out_sf]
    if return_valid:
        return cost_volume, single, torch.stack(out_valid)
    return cost_volume, single


# ----------------------------------------------------------------------------------------------
# closed form -- numpy, explicit gather; dtype selectable
# ----------------------------------------------------------------------------------------------

def _box3(q):
    """3x3 box *sum* with zero padding over the last two axes."""
    p = np.pad(q, [(0, 0)] * (q.ndim - 2) + [(1, 1), (1, 1)])
    h = p[..., :, :-2] + p[..., :, 1:-1] + p[..., :, 2:]
    return h[..., :-2, :] + h[..., 1:-1, :] + h[..., 2:, :]


def _bilinear_zero(img, sx, sy):
    """img (C,H,W); sx, sy (...) source pixel coordinates; taps outside the image contribute 0."""
    C, H, W = img.shape
    x0 = np.floor(sx)
    y0 = np.floor(sy)
    fx = (sx - x0).astype(img.dtype)
    fy = (sy - y0).astype(img.dtype)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    out = np.zeros((C,) + sx.shape, dtype=img.dtype)
    for dy, wy in ((0, 1 - fy), (1, fy)):
        for dx, wx in ((0, 1 - fx), (1, fx)):
            xi, yi = x0 + dx, y0 + dy
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            v = img[:, np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)]
            out += v * (wx * wy * ok)[None]
    return out


def projection_tables(data, use_mono=True, use_stereo=False, dtype=np.float64):
    """proj[b,f] = (K_f . inv(pose_f) . pose_kf)[0:3, 0:4], kinv[b] = inv(K_kf)[0:3, 0:3]  (Appendix C)."""
    frames, intrinsics, poses = collect_frames(data, use_mono, use_stereo)
    B = data["keyframe"].shape[0]
    proj = np.zeros((B, len(frames), 3, 4), dtype=dtype)
    kinv = np.zeros((B, 3, 3), dtype=dtype)
    for b in range(B):
        kinv[b] = np.linalg.inv(data["keyframe_intrinsics"][b].numpy().astype(dtype))[:3, :3]
        for f in range(len(frames)):
            T = np.linalg.inv(poses[f][b].numpy().astype(dtype)) @ data["keyframe_pose"][b].numpy().astype(dtype)
            proj[b, f] = (intrinsics[f][b].numpy().astype(dtype) @ T)[:3, :]
    return proj, kinv


def cost_volume_closed_form(data, inv_depth_min=0.33, inv_depth_max=0.0025, steps=32, use_mono=True,
                            use_stereo=False, alpha=ALPHA, channel_weights=CHANNEL_WEIGHTS, dtype=np.float32):
    """Direct evaluation of the Appendix-C formulas.  Returns (cv, [sfcv_f], valid (B,F,H,W), sad (B,F,D,H,W))."""
    frames, _, _ = collect_frames(data, use_mono, use_stereo)
    key = data["keyframe"].numpy().astype(dtype)
    B, C, H, W = key.shape
    nF, D = len(frames), int(steps)
    proj, kinv = projection_tables(data, use_mono, use_stereo, dtype=np.float64)
    z = (1.0 / np.linspace(float(inv_depth_max), float(inv_depth_min), D, dtype=np.float64))
    if dtype == np.float32:
        z = plane_depths(inv_depth_min, inv_depth_max, D).numpy().astype(np.float64)
    vv, uu = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    inside = np.zeros((H, W), dtype=bool)
    inside[2:H - 2, 2:W - 2] = True
    cw = np.asarray(channel_weights, dtype=dtype).reshape(1, 3, 1, 1)

    cvs = np.zeros((B, D, H, W), dtype=dtype)
    sfs = np.zeros((nF, B, D, H, W), dtype=dtype)
    valids = np.zeros((B, nF, H, W), dtype=bool)
    sads = np.zeros((B, nF, D, H, W), dtype=dtype)
    for b in range(B):
        ray = np.einsum("ij,jhw->ihw", kinv[b], np.stack([uu, vv, np.ones_like(uu)]))      # (3,H,W)
        Y = key[b] + dtype(0.5)
        mu_y = _box3(Y) / dtype(9)
        s_y = _box3(Y * Y) / dtype(9) - mu_y * mu_y
        num = np.zeros((D, H, W), dtype=dtype)
        den = np.zeros((H, W), dtype=dtype)
        for f in range(nF):
            img = frames[f][b].numpy().astype(dtype)
            P = proj[b, f]
            A = np.einsum("ij,jhw->ihw", P[:, :3], ray)                                    # (3,H,W)
            c = A[None] * z[:, None, None, None] + P[:, 3][None, :, None, None]            # (D,3,H,W)
            c = c.astype(dtype).astype(np.float64) if dtype == np.float32 else c
            px = c[:, 0] / (c[:, 2] + 1e-7)
            py = c[:, 1] / (c[:, 2] + 1e-7)
            gx = np.clip((px / (W - 1) - 0.5) * 2, -2, 2)
            gy = np.clip((py / (H - 1) - 0.5) * 2, -2, 2)
            sx = ((gx + 1) * W - 1) / 2
            sy = ((gy + 1) * H - 1) / 2
            if dtype == np.float32:
                sx, sy = sx.astype(np.float32), sy.astype(np.float32)
            X = _bilinear_zero(img, sx, sy) + dtype(0.5)                                   # (3,D,H,W)
            hit = _bilinear_zero(inside[None].astype(dtype), sx, sy)[0] != 0               # (D,H,W)
            valid = inside & hit.all(axis=0)
            X = np.moveaxis(X, 0, 1)                                                       # (D,3,H,W)
            mu_x = _box3(X) / dtype(9)
            s_x = _box3(X * X) / dtype(9) - mu_x * mu_x
            s_xy = _box3(X * Y[None]) / dtype(9) - mu_x * mu_y[None]
            n_ = (2 * mu_x * mu_y[None] + dtype(SSIM_C1)) * (2 * s_xy + dtype(SSIM_C2))
            d_ = (mu_x * mu_x + (mu_y * mu_y)[None] + dtype(SSIM_C1)) * (s_x + s_y[None] + dtype(SSIM_C2))
            e = np.clip((1 - n_ / d_) / 2, 0, 1)
            sad = _box3((e * cw).sum(axis=1)) / dtype(9)                                   # (D,H,W)
            sads[b, f] = sad
            valids[b, f] = valid
            sfs[f, b] = (1 - 2 * sad) * valid
            spread = np.exp(-dtype(alpha) * (sad - sad.min(axis=0, keepdims=True)) ** 2).sum(axis=0)
            w = (1 - (spread - 1) / dtype(D - 1)) * valid
            num += w[None] * sad
            den += w
        nz = den != 0
        cv = np.zeros((D, H, W), dtype=dtype)
        cv[:, nz] = 1 - 2 * num[:, nz] / den[nz]
        cvs[b] = cv
    return cvs, [sfs[f] for f in range(nF)], valids, sads
