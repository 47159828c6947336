"""Host-side mirror of the reference's CostVolumeModule (model/monorec/monorec_model.py:132-284).

Same constructor arguments, same data_dict keys in and out; the arithmetic runs in the fused sm_100a kernel of
libmonorec_b200.so (csrc/cost_volume.cu) through the C ABI.  No torch fallback.
"""
import os
import time

import torch
from torch import nn

from . import _lib


def _as_f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


class CostVolumeModule(nn.Module):
    """Drop-in for the reference class of the same name (monorec_model.py:132-148 for the ctor).

    Supported configuration = the one every shipped config uses: use_ssim=True (1), patch_size=3,
    sfcv_mult_mask=True, not_center_cv=False; use_mono / use_stereo select the frame lists exactly like the
    reference (:160-167).  Other ablation switches raise NotImplementedError instead of silently running something
    else.
    """

    def __init__(self, use_mono=True, use_stereo=False, use_ssim=True, patch_size=3,
                 channel_weights=(5 / 32, 16 / 32, 11 / 32), alpha=10, not_center_cv=False, sfcv_mult_mask=True):
        super().__init__()
        self.use_mono = use_mono
        self.use_stereo = use_stereo
        self.use_ssim = use_ssim
        self.patch_size = patch_size
        self.border_radius = patch_size // 2 + 1
        self.channel_weights = None if channel_weights is None else tuple(float(c) for c in channel_weights)
        self.alpha = alpha
        self.not_center_cv = not_center_cv
        self.sfcv_mult_mask = sfcv_mult_mask
        # False: bilinear taps straight from global memory instead of TMA-staged shared-memory windows (tests, A/B timing)
        self.tma_windows = os.environ.get("MONOREC_B200_CV_TMA", "1") != "0"
        if not (use_ssim is True or use_ssim == 1) or isinstance(use_ssim, float):
            raise NotImplementedError("monorec_b200: only use_ssim=True is implemented (reference default)")
        if patch_size != 3 or not_center_cv or not sfcv_mult_mask:
            raise NotImplementedError("monorec_b200: only patch_size=3, not_center_cv=False, sfcv_mult_mask=True")

    def _gather(self, data_dict):
        frames, intrinsics, poses = [], [], []
        if self.use_mono:
            frames += list(data_dict["frames"])
            intrinsics += list(data_dict["intrinsics"])
            poses += list(data_dict["poses"])
        if self.use_stereo:
            frames += [data_dict["stereoframe"]]
            intrinsics += [data_dict["stereoframe_intrinsics"]]
            poses += [data_dict["stereoframe_pose"]]
        return frames, intrinsics, poses

    @torch.no_grad()
    def forward(self, data_dict):
        start_time = time.time()
        keyframe = _as_f32c(data_dict["keyframe"])
        if not keyframe.is_cuda:
            raise _lib.MonorecLibraryError("monorec_b200.CostVolumeModule needs CUDA tensors (no CPU fallback)")
        if "cv_depths" in data_dict:
            raise NotImplementedError("monorec_b200: per-pixel cv_depths hypotheses are not implemented")
        lib = _lib.load()
        frames, intrinsics, poses = self._gather(data_dict)
        frames = [_as_f32c(f) for f in frames]
        intrinsics = [_as_f32c(k) for k in intrinsics]
        poses = [_as_f32c(p) for p in poses]
        kpose = _as_f32c(data_dict["keyframe_pose"])
        kK = _as_f32c(data_dict["keyframe_intrinsics"])
        B, C, H, W = keyframe.shape
        F = len(frames)
        if C != 3:
            raise NotImplementedError("monorec_b200: 3-channel images only")
        # .item()-free: the ranges are python floats on the model, mirrored into the dict as 1-element tensors
        lo, hi, D = self._plane_range(data_dict)
        dev = keyframe.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            proj = torch.empty(B, F, 3, 4, device=dev, dtype=torch.float32)
            depths = torch.empty(D, device=dev, dtype=torch.float32)
            cv = torch.empty(B, D, H, W, device=dev, dtype=torch.float32)
            sfcv = torch.empty(F, B, D, H, W, device=dev, dtype=torch.float32)
            _lib.check(lib.mr_projection_tables(kpose.data_ptr(), kK.data_ptr(), _lib.ptr_array(poses),
                                                _lib.ptr_array(intrinsics), B, F, H, W, proj.data_ptr(),
                                                depths.data_ptr(), D, lo, hi, stream), "mr_projection_tables")
            cw = None
            if self.channel_weights is not None:
                cw = (_lib.c_float * 3)(*self.channel_weights)
            else:
                cw = (_lib.c_float * 3)(1 / 3, 1 / 3, 1 / 3)  # monorec_model.py:174-177
            nhwc = data_dict.get("_sfcv_nhwc")   # MonoRecModel: the MaskModule's input buffer [F*B,H,W,D], filled by the kernel
            if nhwc is not None and self.tma_windows and D <= 32 and D % 8 == 0 and tuple(nhwc.shape) == (F * B, H, W, D) \
                    and nhwc.is_contiguous() and nhwc.dtype in (torch.float32, torch.float16):
                _lib.check(lib.mr_cost_volume_fwd_nhwc(keyframe.data_ptr(), _lib.ptr_array(frames), proj.data_ptr(),
                                                       depths.data_ptr(), cv.data_ptr(), sfcv.data_ptr(), nhwc.data_ptr(),
                                                       1 if nhwc.dtype == torch.float16 else 0, B, F, D, H, W,
                                                       float(self.alpha), cw, stream), "mr_cost_volume_fwd_nhwc")
                data_dict["_sfcv_nhwc_filled"] = True
            else:
                fwd = lib.mr_cost_volume_fwd if self.tma_windows else lib.mr_cost_volume_fwd_gather
                _lib.check(fwd(keyframe.data_ptr(), _lib.ptr_array(frames), proj.data_ptr(), depths.data_ptr(), cv.data_ptr(),
                               sfcv.data_ptr(), B, F, D, H, W, float(self.alpha), cw, stream), "mr_cost_volume_fwd")
        data_dict["cost_volume"] = cv
        data_dict["single_frame_cvs"] = [sfcv[f] for f in range(F)]
        # host-side issue time (the reference's number includes its device work only because it synchronises implicitly)
        data_dict["cv_module_time"] = torch.full((1,), time.time() - start_time, device=dev, dtype=torch.float32)
        return data_dict

    @staticmethod
    def _plane_range(data_dict):
        """(inv_depth_lo, inv_depth_hi, D) from the dict (monorec_model.py:184: names are swapped w.r.t. values).

        The model stores python numbers next to the tensors (keys with a leading underscore) so that no device->host
        synchronisation is needed; a dict that only has the reference's tensors falls back to .item().
        """
        if "_cv_range" in data_dict:
            return data_dict["_cv_range"]
        return (float(data_dict["inv_depth_max"][0].item()), float(data_dict["inv_depth_min"][0].item()),
                int(data_dict["cv_depth_steps"][0].item()))

    def create_mask(self, c, height, width, border_radius, device=None):
        """Kept for API parity (monorec_model.py:282-284); the kernel never materialises this mask."""
        mask = torch.zeros(c, 1, height, width, device=device)
        mask[:, :, border_radius:height - border_radius, border_radius:width - border_radius] = 1
        return mask
