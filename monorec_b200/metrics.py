"""Drop-ins for the reference's sparse depth metrics (model/metric_functions/sparse_metrics.py:81-251) and the loader's image
normalisation (data_loader/kitti_odometry_dataset.py:121-132), computed on the device by libmonorec_b200.so.

The reference's evaluater (evaluater/evaluater.py:78-112) calls the seven `*_sparse_metric(data_dict, roi, max_distance)`
functions one after the other, each a dozen elementwise torch kernels and a few reductions.  Here all seven come out of ONE
fused pass (`mr_sparse_metrics`); the functions below keep the reference's names and signatures and share that pass through a
small per-data_dict cache, so `model.metric` can be pointed at this module unchanged.  No CPU fallback.
"""
import ctypes

import torch

from . import _lib

NAMES = ("a1", "a2", "a3", "rmse", "rmse_log", "abs_rel", "sq_rel")


def sparse_metrics(data_dict, roi=None, max_distance=None, pred_all_valid=True, use_cvmask=False):
    """-> device tensor [7]: a1, a2, a3, rmse, rmse_log, abs_rel, sq_rel (no host synchronisation)."""
    pred, gt = data_dict["result"], data_dict["target"]
    if not pred.is_cuda:
        raise _lib.MonorecLibraryError("monorec_b200.metrics needs CUDA tensors (no CPU fallback)")
    key = (id(pred), pred._version, id(gt), gt._version, None if roi is None else tuple(int(v) for v in roi),
           None if max_distance is None else float(max_distance), bool(pred_all_valid), bool(use_cvmask))
    cache = data_dict.get("_mr_metrics_cache")
    if cache is not None and cache[0] == key:
        return cache[1]
    lib = _lib.load()
    pred = pred.to(torch.float32).contiguous()
    gt = gt.to(device=pred.device, dtype=torch.float32).contiguous()
    B, _, H, W = pred.shape
    mv = None
    if use_cvmask:
        mv = data_dict["mvobj_mask"].to(device=pred.device, dtype=torch.float32).contiguous()
    out = torch.empty(7, device=pred.device, dtype=torch.float32)
    ws_bytes = lib.mr_sparse_metrics_workspace(B)
    ws = torch.empty(ws_bytes // 8, device=pred.device, dtype=torch.float64)
    roi_c = None if roi is None else (ctypes.c_int * 4)(*[int(v) for v in roi])
    with torch.cuda.device(pred.device):
        _lib.check(lib.mr_sparse_metrics(pred.data_ptr(), gt.data_ptr(), None if mv is None else mv.data_ptr(), B, H, W, roi_c,
                                         float(max_distance) if max_distance else 0.0, 1 if pred_all_valid else 0,
                                         out.data_ptr(), ws.data_ptr(), ws_bytes,
                                         torch.cuda.current_stream(pred.device).cuda_stream), "mr_sparse_metrics")
    data_dict["_mr_metrics_cache"] = (key, out)
    return out


def _make(index, **fixed):
    def metric(data_dict, roi=None, max_distance=None, pred_all_valid=True, use_cvmask=False):
        kw = dict(pred_all_valid=pred_all_valid, use_cvmask=use_cvmask)
        kw.update(fixed)
        return sparse_metrics(data_dict, roi, max_distance, **kw)[index]
    return metric


for _i, _n in enumerate(NAMES):
    globals()[f"{_n}_sparse_metric"] = _make(_i)                                        # sparse_metrics.py:81-156
    globals()[f"{_n}_sparse_onlyvalid_metric"] = _make(_i, pred_all_valid=False)        # :159-184
    globals()[f"{_n}_sparse_onlydynamic_metric"] = _make(_i, use_cvmask=True)           # :187-212


def images_u8_to_f32(images_u8, crop_box=None):
    """uint8 HWC images [B,Hs,Ws,3] on the device -> float CHW [B,3,H,W] = u / 255 - 0.5, optionally cropped to the PIL-style
    box (left, upper, right, lower) -- kitti_odometry_dataset.py:121-132 without the resize."""
    if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[3] != 3:
        raise _lib.MonorecLibraryError("images_u8_to_f32 expects a CUDA uint8 tensor [B,H,W,3]")
    lib = _lib.load()
    x = images_u8.contiguous()
    B, Hs, Ws, _ = x.shape
    left, top, right, bottom = (0, 0, Ws, Hs) if crop_box is None else [int(v) for v in crop_box]
    H, W = bottom - top, right - left
    out = torch.empty(B, 3, H, W, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.mr_images_u8_to_f32(x.data_ptr(), out.data_ptr(), B, Hs, Ws, top, left, H, W,
                                           torch.cuda.current_stream(x.device).cuda_stream), "mr_images_u8_to_f32")
    return out
