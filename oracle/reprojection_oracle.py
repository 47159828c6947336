"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference's photometric reprojection loss.

Only `tests/` may import this file, and only as the checker.  The product path (monorec_b200/losses.py ->
libmonorec_b200.so) never imports anything from `oracle/`.

Reference being restated: model/loss_functions/common_losses.py:10-13 (compute_errors) and :16-114 (reprojection_loss) for
the argument sets the reference's losses actually use (model/loss_functions/monorec_loss.py:185-188, :264-265, :355, :361):
error_function=compute_errors, combine_frames="min", mono_auto=False, automasking in {False, True}, border in {0, 3},
use_mono / use_stereo.  Helpers: model/layers.py:43-61 (Backprojection), :63-71 (point_projection), :79-89 (GaussianAverage),
:91-139 (SSIM with zero padding, Gaussian window, comp_mode), utils/util.py:110-118 (mask_mean), :130-132 (create_mask).

Written with the reference's own primitives (torch.inverse, matmul, F.grid_sample, F.conv2d) in the reference's order, so on
the CPU it reproduces the reference bit for bit, and its gradient is torch autograd of that forward -- which is what
`loss.backward()` computes in trainer/monorec_trainer.py:143-145.

Parity pin: tests/golden/reprojection.npz, written by tests/golden/make_golden.py --only-reprojection from the UNMODIFIED
reference function (errors and d(sum_w errors)/d depth_prediction for three argument sets); tests/test_reprojection.py
checks this restatement against it on the CPU and the CUDA path against both.
"""
import torch
import torch.nn.functional as F

GAUSS = ((0.0947, 0.1183, 0.0947), (0.1183, 0.1478, 0.1183), (0.0947, 0.1183, 0.0947))   # layers.py:82-85
SSIM_C1 = 0.01 ** 2                                                                        # layers.py:116
SSIM_C2 = 0.03 ** 2                                                                        # layers.py:117


def _gauss(x):
    """layers.py:87-89: depthwise 3x3 correlation with the fixed window, no padding."""
    k = torch.tensor(GAUSS, dtype=x.dtype).repeat(x.shape[1], 1, 1, 1)
    return F.conv2d(x, k, padding=0, groups=x.shape[1])


def _ssim_comp(x, y):
    """layers.py:119-139 with pad_reflection=False, gaussian_average=True, comp_mode=True."""
    x = F.pad(x, (1, 1, 1, 1))
    y = F.pad(y, (1, 1, 1, 1))
    mu_x, mu_y = _gauss(x), _gauss(y)
    mu_x_sq, mu_y_sq, mu_x_y = mu_x ** 2, mu_y ** 2, mu_x * mu_y
    sigma_x = _gauss(x ** 2) - mu_x_sq
    sigma_y = _gauss(y ** 2) - mu_y_sq
    sigma_xy = _gauss(x * y) - mu_x_y
    n = (2 * mu_x_y + SSIM_C1) * (2 * sigma_xy + SSIM_C2)
    d = (mu_x_sq + mu_y_sq + SSIM_C1) * (sigma_x + sigma_y + SSIM_C2)
    return torch.clamp(1 - n / d, 0, 1) / 2


def compute_errors(img0, img1):
    """common_losses.py:10-13."""
    return .85 * torch.mean(_ssim_comp(img0, img1), dim=1) + .15 * torch.mean(torch.abs(img0 - img1), dim=1)


def _pixel_rows(height, width, dtype):
    # layers.py:49-54: rows [x; y; 1], row-major over (y, x)
    ys, xs = torch.meshgrid(torch.arange(height, dtype=dtype), torch.arange(width, dtype=dtype), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, dtype=dtype)], 0)


def _grid(depth_prediction, kf_K, K, T, height, width):
    """Backprojection.forward (layers.py:56-61) with depth = 1 / depth_prediction, then point_projection (:63-71)."""
    B = depth_prediction.shape[0]
    coord = _pixel_rows(height, width, depth_prediction.dtype).unsqueeze(0).repeat(B, 1, 1)
    cam_norm = torch.matmul(torch.inverse(kf_K)[:, :3, :3], coord)
    cam = (1 / depth_prediction).view(B, 1, -1) * cam_norm
    cam = torch.cat([cam, torch.ones(B, 1, height * width, dtype=cam.dtype)], 1)
    c = torch.matmul(torch.matmul(K, T)[:, :3, :], cam)
    img = c[:, :2, :] / (c[:, 2:3, :] + 1e-7)
    img = torch.stack([img[:, 0, :] / (width - 1), img[:, 1, :] / (height - 1)], 1)
    img = (img - 0.5) * 2
    return img.view(B, 2, height, width).permute(0, 2, 3, 1)


def collect(data, use_mono=True, use_stereo=False):
    """common_losses.py:23-34."""
    frames, poses, intrinsics = [], [], []
    if use_mono:
        frames += list(data["frames"]); poses += list(data["poses"]); intrinsics += list(data["intrinsics"])
    if use_stereo:
        frames.append(data["stereoframe"]); poses.append(data["stereoframe_pose"]); intrinsics.append(data["stereoframe_intrinsics"])
    return frames, poses, intrinsics


def reprojection_errors(depth_prediction, data, automasking=False, use_mono=True, use_stereo=False, border=0):
    """-> (errors [B,H,W] with +inf where no source frame gives a usable sample, winner [B,H,W] int64 (-1: none)).

    common_losses.py:16-114 for reduce=False, combine_frames="min".  Differentiable w.r.t. depth_prediction [B,1,H,W]."""
    key = data["keyframe"]
    B, C, H, W = key.shape
    frames, poses, intrinsics = collect(data, use_mono, use_stereo)
    Fn = len(frames)
    reproj, wmasks = [], []
    for frame, pose, K in zip(frames, poses, intrinsics):                                       # :49-54
        grid = _grid(depth_prediction, data["keyframe_intrinsics"], K, torch.inverse(pose) @ data["keyframe_pose"], H, W)
        reproj.append(F.grid_sample(frame + 1.5, grid, padding_mode="zeros", align_corners=False))
        if border > 0:
            m = F.pad(torch.ones(B, 1, H - 2 * border, W - 2 * border, dtype=key.dtype), [border] * 4)   # util.py:130-132
            wmasks.append(F.grid_sample(m, grid, padding_mode="zeros", align_corners=False))
    reproj = torch.stack(reproj, dim=1).view(B * Fn, C, H, W)                                  # :56
    mask = reproj[:, 0, :, :] == 0                                                             # :57
    reproj = reproj - 1.0                                                                      # :58
    if border > 0:
        mask = ~(torch.stack(wmasks, dim=1).view(B * Fn, H, W) > .5)                           # :60-61
    key_exp = (key + .5).unsqueeze(1).expand(-1, Fn, -1, -1, -1).reshape(B * Fn, C, H, W)      # :63
    errors = compute_errors(reproj, key_exp).view(B, Fn, H, W)                                 # :73-75
    mask = mask.view(B, Fn, H, W)
    errors = torch.where(mask, torch.full_like(errors, float("inf")), errors)                  # :78
    if automasking:                                                                            # :80-83
        stacked = torch.stack(frames, dim=1).view(B * Fn, C, H, W) + .5
        nowarp = compute_errors(stacked, key_exp).view(B, Fn, H, W)
        errors = torch.where(nowarp < errors, torch.full_like(errors, float("inf")), errors)
    best, idx = torch.min(errors, dim=1)                                                       # :93-95
    idx = torch.where(torch.isinf(best), torch.full_like(idx, -1), idx)
    return best, idx


def mask_mean(t, m):
    """utils/util.py:110-118 over all dimensions."""
    t = torch.where(m, torch.zeros_like(t), t)
    return torch.sum(t) / (t.numel() - torch.sum(m.to(torch.float)))


def reprojection_loss(depth_prediction, data, reduce=True, **kw):
    errors, _ = reprojection_errors(depth_prediction, data, **kw)
    if reduce:                                                                                 # :110-111
        return mask_mean(errors, torch.isinf(errors))
    return errors


# ---- second, independent restatement: closed form in numpy (explicit homography, bilinear gather, analytic gradient) ----------
def reprojection_closed_form(depth_prediction, data, grad_errors=None, automasking=False, use_mono=True, use_stereo=False,
                             border=0, dtype="float64", kink_eps=None):
    """-> (errors [B,H,W], winner [B,H,W], grad [B,1,H,W] or None).  Shares no primitive with `reprojection_errors`; the gradient
    is derived by hand (the derivation the CUDA kernel uses):

      R = n / d of layers.py:133-134 for one window p, with Gaussian weights g(p - q) over its 9 pixels q:
      d R / d x(q) = g(p - q) (alpha + beta y(q) + gamma x(q)),
        alpha = 2 mu_y (A2 - A1) / d - 2 (R / d) mu_x (B2 - B1),  beta = 2 A1 / d,  gamma = -2 (R / d) B1,
        A1 = 2 mu_x mu_y + C1, A2 = 2 sigma_xy + C2, B1 = mu_x^2 + mu_y^2 + C1, B2 = sigma_x + sigma_y + C2;
      d x(q) / d inv_depth(q) through the bilinear taps inside the image and s = c_xy / c_z, c = a / inv_depth + t.
    """
    import numpy as np
    ft = np.dtype(dtype).type
    key = data["keyframe"].double().numpy().astype(dtype)
    B, C, H, W = key.shape
    frames, poses, intrinsics = collect(data, use_mono, use_stereo)
    invd = depth_prediction.detach().double().numpy().astype(dtype)[:, 0]
    y = np.pad(key + ft(0.5), ((0, 0), (0, 0), (1, 1), (1, 1)))
    g = np.array(GAUSS, dtype=dtype)
    vv, uu = np.meshgrid(np.arange(H, dtype=dtype), np.arange(W, dtype=dtype), indexing="ij")
    kinv = np.linalg.inv(data["keyframe_intrinsics"].double().numpy())[:, :3, :3]

    def gsum(t):          # [B,C,H+2,W+2] -> [B,C,H,W]
        out = np.zeros(t.shape[:2] + (H, W), dtype=dtype)
        for dy in range(3):
            for dx in range(3):
                out += g[dy, dx] * t[:, :, dy:dy + H, dx:dx + W]
        return out

    def stats(xp):
        mx, my = gsum(xp), gsum(y)
        sxx, syy, sxy = gsum(xp * xp), gsum(y * y), gsum(xp * y)
        A1, A2 = 2 * mx * my + SSIM_C1, 2 * (sxy - mx * my) + SSIM_C2
        B1, B2 = mx * mx + my * my + SSIM_C1, (sxx - mx * mx) + (syy - my * my) + SSIM_C2
        return mx, my, A1, A2, B1, B2

    def errors_of(x):     # x: [B,3,H,W]
        xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
        mx, my, A1, A2, B1, B2 = stats(xp)
        ssim = np.clip(1 - A1 * A2 / (B1 * B2), 0, 1) / 2
        return ft(.85) * ssim.mean(1) + ft(.15) * np.abs(x - (key + ft(0.5))).mean(1)

    per_frame, warped, derivs, frac_kinks = [], [], [], []
    for frame, pose, K in zip(frames, poses, intrinsics):
        P = (K.double().numpy() @ (np.linalg.inv(pose.double().numpy()) @ data["keyframe_pose"].double().numpy()))[:, :3, :]
        M = P[:, :, :3] @ kinv                                             # [B,3,3]
        t = P[:, :, 3].copy()
        t[:, 2] += 1e-7                                                    # layers.py:66
        img = frame.double().numpy().astype(dtype)
        a = [(M[:, i, 0, None, None] * uu + M[:, i, 1, None, None] * vv + M[:, i, 2, None, None]).astype(dtype) for i in range(3)]
        z = 1 / invd
        cx, cy, cz = a[0] * z + t[:, 0, None, None], a[1] * z + t[:, 1, None, None], a[2] * z + t[:, 2, None, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            ux, uy = cx / cz, cy / cz                                      # pixel units of point_projection
            sx, sy = ux * W / (W - 1) - 0.5, uy * H / (H - 1) - 0.5        # grid_sample(align_corners=False) of (u/(W-1) - .5) * 2
        sx = np.where(np.isfinite(sx), sx, -1e9); sy = np.where(np.isfinite(sy), sy, -1e9)
        x0, y0 = np.floor(sx), np.floor(sy)
        fx, fy = sx - x0, sy - y0
        x0i, y0i = x0.astype(np.int64), y0.astype(np.int64)
        bi = np.arange(B)[:, None, None]
        taps, tin = {}, {}
        for name, (oy, ox) in {"nw": (0, 0), "ne": (0, 1), "sw": (1, 0), "se": (1, 1)}.items():
            xi, yi = x0i + ox, y0i + oy
            inb = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            xc, yc = np.clip(xi, 0, W - 1), np.clip(yi, 0, H - 1)
            taps[name] = np.where(inb[:, None], img[bi[:, None], np.arange(3)[None, :, None, None], yc[:, None], xc[:, None]] + ft(1.5), 0)
            tin[name] = (xi >= border) & (xi < W - border) & (yi >= border) & (yi < H - border) & inb
        wnw, wne, wsw, wse = (1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy
        raw = taps["nw"] * wnw[:, None] + taps["ne"] * wne[:, None] + taps["sw"] * wsw[:, None] + taps["se"] * wse[:, None]
        if border > 0:
            masked = ~((tin["nw"] * wnw + tin["ne"] * wne + tin["sw"] * wsw + tin["se"] * wse) > 0.5)
        else:
            masked = raw[:, 0] == 0
        x = raw - 1
        e = errors_of(x)
        e = np.where(masked, np.inf, e)
        if automasking:
            e0 = errors_of(img + ft(0.5))
            e = np.where(e0 < e, np.inf, e)
        per_frame.append(e)
        warped.append(x)
        gxs = (taps["ne"] - taps["nw"]) * (1 - fy)[:, None] + (taps["se"] - taps["sw"]) * fy[:, None]
        gys = (taps["sw"] - taps["nw"]) * (1 - fx)[:, None] + (taps["se"] - taps["ne"]) * fx[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            dz = -z * z
            dsx = dz * (a[0] - ux * a[2]) / cz * W / (W - 1)
            dsy = dz * (a[1] - uy * a[2]) / cz * H / (H - 1)
        derivs.append((gxs, gys, np.nan_to_num(dsx), np.nan_to_num(dsy)))
        eps = 0 if kink_eps is None else kink_eps
        frac_kinks.append((np.minimum(fx, 1 - fx) < eps) | (np.minimum(fy, 1 - fy) < eps))
    E = np.stack(per_frame, 1)
    winner = E.argmin(1)
    best = E.min(1)
    winner = np.where(np.isinf(best), -1, winner)
    if grad_errors is None:
        return best, winner, None
    ge = np.asarray(grad_errors, dtype=dtype)
    grad = np.zeros((B, H, W), dtype=dtype)
    kink = np.zeros((B, H, W), dtype=bool)
    for f, (x, (gxs, gys, dsx, dsy)) in enumerate(zip(warped, derivs)):
        gsel = np.where(winner == f, ge, 0)[:, None]                       # upstream gradient of the windows this frame wins
        xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
        mx, my, A1, A2, B1, B2 = stats(xp)
        d = B1 * B2
        R = A1 * A2 / d
        inside = ((1 - R) >= 0) & ((1 - R) <= 1)
        sc = np.where(inside, -0.5 * 0.85 / 3, 0) * gsel
        al = sc * (2 * my * (A2 - A1) / d - 2 * R / d * mx * (B2 - B1))
        be = sc * (2 * A1 / d)
        ga = sc * (-2 * R / d * B1)
        pad = lambda t: np.pad(t, ((0, 0), (0, 0), (1, 1), (1, 1)))       # noqa: E731  windows outside the image do not exist
        yq = key + ft(0.5)
        Gx = gsum(pad(al)) + gsum(pad(be)) * yq + gsum(pad(ga)) * x + gsel * (0.15 / 3) * np.sign(x - yq)
        grad += (Gx * (gxs * dsx[:, None] + gys * dsy[:, None])).sum(1)
        if kink_eps is not None:
            won = winner == f
            kw = won & ((np.abs(R) < kink_eps) | (np.abs(1 - R) < kink_eps)).any(1)          # windows at a clamp bound
            kp = np.pad(kw, ((0, 0), (1, 1), (1, 1)))
            reach = np.zeros_like(kw)
            for dy in range(3):
                for dx in range(3):
                    reach |= kp[:, dy:dy + H, dx:dx + W]
            kink |= reach | (won & (np.abs(x - yq) < kink_eps).any(1))
            near = np.pad(won, ((0, 0), (1, 1), (1, 1)))
            used = np.zeros_like(won)
            for dy in range(3):
                for dx in range(3):
                    used |= near[:, dy:dy + H, dx:dx + W]
            kink |= used & frac_kinks[f]
    if kink_eps is not None:
        return best, winner, grad[:, None], kink
    return best, winner, grad[:, None]
