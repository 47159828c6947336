"""Host side of the convolution engine: ctypes descriptor, weight packing and layer-level helpers.

Activations inside the engine are NHWC fp32 torch tensors; every helper launches kernels of libmonorec_b200.so on the
current CUDA stream through the C ABI (include/monorec_b200.h: mr_conv_desc).  Nothing here computes on the CPU and
nothing falls back to torch ops.
"""
import ctypes
import math
from ctypes import c_float, c_int, c_void_p

import torch

from . import _lib

MAX_SRC = 3
ACT_NONE, ACT_LEAKY, ACT_SIGMOID, ACT_ABSTANH = 0, 1, 2, 3
LEAKY_SLOPE = 0.1  # model/layers.py:290, 318, 381


class ConvDesc(ctypes.Structure):
    """Mirror of `struct mr_conv_desc` (include/monorec_b200.h)."""
    _fields_ = [("n_src", c_int), ("src", c_void_p * MAX_SRC), ("src_c", c_int * MAX_SRC),
                ("B", c_int), ("Hs", c_int), ("Ws", c_int), ("upsample2", c_int),
                ("kh", c_int), ("kw", c_int), ("sy", c_int), ("sx", c_int), ("pad_t", c_int), ("pad_l", c_int),
                ("Ho", c_int), ("Wo", c_int), ("Cout", c_int),
                ("weight", c_void_p), ("bias", c_void_p), ("dst", c_void_p),
                ("dst_H", c_int), ("dst_W", c_int), ("dst_c", c_int), ("dst_coff", c_int),
                ("oy_step", c_int), ("ox_step", c_int), ("oy_off", c_int), ("ox_off", c_int),
                ("act", c_int), ("act_a", c_float), ("act_b", c_float)]


def same_pad_before(n, k, s):
    """Leading zero padding of PadSameConv2d (model/layers.py:249-251); the trailing part is implicit (zero fill)."""
    total = s * (math.ceil(n / s) - 1) + k - n
    return total // 2


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def pack_conv_weight(w):
    """nn.Conv2d weight (Cout, Cin, kh, kw) -> [kh][kw][Cin][Cout] contiguous fp32."""
    return w.detach().to(torch.float32).permute(2, 3, 1, 0).contiguous()


def pack_convT_k4s2(w):
    """nn.ConvTranspose2d(k=4, s=2) weight (Cin, Cout, 4, 4) -> four sub-pixel 2x2 kernels [py][px] -> [2][2][Cin][Cout].

    With the reference's centre crop of one pixel (model/layers.py:269-286, oversize = -2) output pixel (Y, X) of the
    cropped 2H x 2W map receives input rows i with 2 i + ky = Y + 1:
        Y even (py = 0): (i, ky) = (Y/2 - 1, 3), (Y/2, 1)          Y odd (py = 1): (i, ky) = ((Y-1)/2, 2), ((Y+1)/2, 0)
    i.e. a 2-tap filter along each axis on the input grid with taps ordered by increasing i.
    """
    w = w.detach().to(torch.float32)
    taps = {0: (3, 1), 1: (2, 0)}
    out = {}
    for py in (0, 1):
        for px in (0, 1):
            sub = w[:, :, list(taps[py]), :][:, :, :, list(taps[px])]      # (Cin, Cout, 2, 2)
            out[(py, px)] = sub.permute(2, 3, 0, 1).contiguous()           # [2][2][Cin][Cout]
    return out


def conv2d(srcs, weight, bias, kh, kw, stride=(1, 1), act=ACT_NONE, act_a=0.0, act_b=1.0, upsample2=False,
           out=None, out_coff=0, pad=None, out_hw=None, out_step=(1, 1), out_off=(0, 0)):
    """One fused convolution launch.

    srcs: list of NHWC tensors [B, Hs, Ws, C_i] (concatenated along C in this order).  weight: packed [kh][kw][Cin][Cout].
    Returns the NHWC output tensor (allocated unless `out` is given; then the channel slice at `out_coff` is written).
    """
    lib = _lib.load()
    x0 = srcs[0]
    B, Hs, Ws, _ = x0.shape
    Hv, Wv = (2 * Hs, 2 * Ws) if upsample2 else (Hs, Ws)
    sy, sx = stride
    Cout = weight.shape[-1]
    if pad is None:
        pad = (same_pad_before(Hv, kh, sy), same_pad_before(Wv, kw, sx))
    if out_hw is None:
        out_hw = (math.ceil(Hv / sy), math.ceil(Wv / sx))
    Ho, Wo = out_hw
    if out is None:
        out = torch.empty(B, Ho * out_step[0], Wo * out_step[1], Cout, device=x0.device, dtype=torch.float32)
    d = ConvDesc()
    d.n_src = len(srcs)
    cin = 0
    for i, s in enumerate(srcs):
        assert s.is_cuda and s.dtype == torch.float32 and s.is_contiguous(), "conv sources must be contiguous fp32 CUDA"
        assert s.shape[:3] == x0.shape[:3], "concatenated sources must share B, H, W"
        d.src[i] = s.data_ptr()
        d.src_c[i] = s.shape[3]
        cin += s.shape[3]
    assert weight.shape == (kh, kw, cin, Cout), f"packed weight {tuple(weight.shape)} != {(kh, kw, cin, Cout)}"
    d.B, d.Hs, d.Ws, d.upsample2 = B, Hs, Ws, int(upsample2)
    d.kh, d.kw, d.sy, d.sx, d.pad_t, d.pad_l = kh, kw, sy, sx, pad[0], pad[1]
    d.Ho, d.Wo, d.Cout = Ho, Wo, Cout
    d.weight = weight.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.dst = out.data_ptr()
    d.dst_H, d.dst_W, d.dst_c, d.dst_coff = out.shape[1], out.shape[2], out.shape[3], out_coff
    d.oy_step, d.ox_step, d.oy_off, d.ox_off = out_step[0], out_step[1], out_off[0], out_off[1]
    d.act, d.act_a, d.act_b = act, act_a, act_b
    with torch.cuda.device(x0.device):
        _lib.check(lib.mr_conv2d_nhwc(ctypes.byref(d), _stream(x0)), "mr_conv2d_nhwc")
    return out


def conv_transpose_k4s2_crop(srcs, sub_weights, bias, act=ACT_LEAKY, act_a=LEAKY_SLOPE):
    """Refine (model/layers.py:380-400): ConvTranspose2d(k4, s2) + LeakyReLU + centre crop, as 4 sub-pixel 2x2 convs."""
    x0 = srcs[0]
    B, Hs, Ws, _ = x0.shape
    Cout = sub_weights[(0, 0)].shape[-1]
    out = torch.empty(B, 2 * Hs, 2 * Ws, Cout, device=x0.device, dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            conv2d(srcs, sub_weights[(py, px)], bias, 2, 2, act=act, act_a=act_a, out=out,
                   pad=(1 - py, 1 - px), out_hw=(Hs, Ws), out_step=(2, 2), out_off=(py, px))
    return out


def nchw_to_nhwc(x, out=None, out_coff=0, one_minus=None):
    """(B,C,H,W) -> NHWC (optionally into a channel slice of `out`, optionally scaled by (1 - one_minus[b,0,h,w]))."""
    lib = _lib.load()
    x = x.contiguous()
    B, C, H, W = x.shape
    if out is None:
        out = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32)
    om = one_minus.contiguous().data_ptr() if one_minus is not None else None
    with torch.cuda.device(x.device):
        _lib.check(lib.mr_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), B, C, H, W, out.shape[3], out_coff, om, _stream(x)),
                   "mr_nchw_to_nhwc")
    return out


def maxpool2(x):
    lib = _lib.load()
    B, H, W, C = x.shape
    out = torch.empty(B, H // 2, W // 2, C, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.mr_maxpool2_nhwc(x.data_ptr(), out.data_ptr(), B, H, W, C, _stream(x)), "mr_maxpool2_nhwc")
    return out


def max_over_frames(x, frames):
    """x: [frames*B, ...] -> [B, ...] element-wise max over the leading frame axis."""
    if frames == 1:
        return x
    lib = _lib.load()
    B = x.shape[0] // frames
    out = torch.empty((B,) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.mr_max_over_frames(x.data_ptr(), out.data_ptr(), frames, out.numel(), _stream(x)),
                   "mr_max_over_frames")
    return out


def mask_volume(volume, mask):
    """cost_volume * (1 - cv_mask) on NCHW tensors (model/monorec/monorec_model.py:713)."""
    lib = _lib.load()
    volume = volume.contiguous()
    mask = mask.to(torch.float32).contiguous()
    B, D, H, W = volume.shape
    out = torch.empty_like(volume)
    with torch.cuda.device(volume.device):
        _lib.check(lib.mr_mask_volume(volume.data_ptr(), mask.data_ptr(), out.data_ptr(), B, D, H * W, _stream(volume)),
                   "mr_mask_volume")
    return out
