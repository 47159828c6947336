"""Parity of the convolution engine / MaskModule / DepthModule / MonoRecModel (through the C ABI) on a B200."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-4   # fp32 CUDA-core path vs fp32 CPU reference: accumulation-order noise only (relative to max|ref|)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


@pytest.mark.parametrize("cin,cout,kh,kw,sy,sx,H,W", [
    (32, 32, 3, 3, 1, 1, 16, 32), (35, 48, 7, 1, 1, 1, 24, 40), (48, 64, 7, 1, 2, 1, 24, 40), (64, 64, 1, 7, 1, 2, 12, 40),
    (64, 128, 5, 1, 2, 1, 20, 24), (128, 128, 1, 5, 1, 2, 10, 24), (192, 256, 3, 1, 2, 1, 18, 16), (24, 1, 3, 3, 1, 1, 16, 32),
    (48, 1, 1, 1, 1, 1, 8, 16), (32, 24, 3, 3, 1, 1, 9, 21), (96, 96, 3, 3, 1, 1, 7, 13), (3, 5, 3, 3, 1, 1, 5, 5)])
def test_conv_same_padding_matches_torch(cin, cout, kh, kw, sy, sx, H, W):
    from monorec_b200 import conv as C
    from oracle.convnet_oracle import conv_same
    g = torch.Generator().manual_seed(cin * 131 + cout)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.leaky_relu(conv_same(x, w, b, (sy, sx)), 0.1)
    out = C.conv2d([_nhwc(x).to(DEV)], C.pack_conv_weight(w).to(DEV), b.to(DEV), kh, kw, stride=(sy, sx),
                   act=C.ACT_LEAKY, act_a=0.1)
    assert _rel(_nchw(out.cpu()), ref) < TOL


def test_concat_sources_upconv_and_refine():
    from monorec_b200 import conv as C
    from oracle import convnet_oracle as CO
    g = torch.Generator().manual_seed(3)
    a, b_, c = torch.randn(2, 96, 8, 16, generator=g), torch.randn(2, 128, 8, 16, generator=g), torch.randn(2, 35, 8, 16, generator=g)
    cat = torch.cat([a, b_, c], 1)
    srcs = [_nhwc(t).to(DEV) for t in (a, b_, c)]
    # 3x3 over a 3-way concatenation
    w = torch.randn(64, 259, 3, 3, generator=g) / 48
    bias = torch.randn(64, generator=g)
    ref = CO.lrelu(CO.conv_same(cat, w, bias))
    out = C.conv2d(srcs, C.pack_conv_weight(w).to(DEV), bias.to(DEV), 3, 3, act=C.ACT_LEAKY, act_a=0.1)
    assert _rel(_nchw(out.cpu()), ref) < TOL
    # Upconv (nearest x2 + pad(0,1,0,1) + 2x2 conv, no activation)
    sd = {"u.conv.weight": torch.randn(96, 259, 2, 2, generator=g) / 32, "u.conv.bias": torch.randn(96, generator=g)}
    ref = CO.upconv(sd, "u", cat)
    out = C.conv2d(srcs, C.pack_conv_weight(sd["u.conv.weight"]).to(DEV), sd["u.conv.bias"].to(DEV), 2, 2, upsample2=True)
    assert out.shape[1:3] == (16, 32) and _rel(_nchw(out.cpu()), ref) < TOL
    # Refine (ConvTranspose2d k4 s2 + LReLU + crop)
    sd = {"r.conv2d_t.weight": torch.randn(259, 48, 4, 4, generator=g) / 32, "r.conv2d_t.bias": torch.randn(48, generator=g)}
    ref = CO.refine(sd, "r", cat)
    sub = {k: v.to(DEV) for k, v in C.pack_convT_k4s2(sd["r.conv2d_t.weight"]).items()}
    out = C.conv_transpose_k4s2_crop(srcs, sub, sd["r.conv2d_t.bias"].to(DEV))
    assert out.shape[1:3] == (16, 32) and _rel(_nchw(out.cpu()), ref) < TOL


def test_small_ops():
    from monorec_b200 import conv as C
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 32, 12, 20, generator=g)
    assert torch.equal(C.nchw_to_nhwc(x.to(DEV)).cpu(), _nhwc(x))
    m = torch.rand(3, 1, 12, 20, generator=g)
    buf = torch.zeros(3, 12, 20, 35, device=DEV)
    C.nchw_to_nhwc(x.to(DEV), out=buf, out_coff=0, one_minus=m.to(DEV))
    C.nchw_to_nhwc(x[:, :3].contiguous().to(DEV), out=buf, out_coff=32)
    assert torch.allclose(buf.cpu()[..., :32], _nhwc(x * (1 - m)), atol=1e-7) and torch.equal(buf.cpu()[..., 32:], _nhwc(x[:, :3]))
    assert torch.equal(_nchw(C.maxpool2(_nhwc(x).to(DEV)).cpu()), F.max_pool2d(x, 2))
    xs = _nhwc(x).to(DEV)
    assert torch.equal(C.max_over_frames(xs, 3).cpu(), _nhwc(x).max(0, keepdim=True)[0])
    assert torch.allclose(C.mask_volume(x.to(DEV), m.to(DEV)).cpu(), x * (1 - m), atol=1e-7)


def _model_and_sd(gain, seed=7):
    from monorec_b200.model import MonoRecModel
    from monorec_b200.synthetic import seeded_state_dict
    model = MonoRecModel()
    sd = seeded_state_dict(model, seed=seed, gain=gain)
    model.load_state_dict(sd)
    return model.to(DEV).eval(), sd


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_pool_and_frame_max_is_the_two_ops(dtype):
    """One pass over an encoder level's output = nn.MaxPool2d(2) per frame + element-wise max over the frames, bit for bit."""
    from monorec_b200 import conv as C
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3 * 2, 12, 20, 48, generator=g).to(DEV, dtype)                       # 3 frames x batch 2
    pooled, fmax = C.pool_and_frame_max(x, 3)
    ref_pool = F.max_pool2d(x.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).to(dtype)
    ref_max = x.view(3, 2, 12, 20, 48).amax(0)
    assert torch.equal(pooled, ref_pool) and torch.equal(fmax, ref_max)
    assert torch.equal(pooled, C.maxpool2(x)) and torch.equal(fmax, C.max_over_frames(x, 3))


@pytest.mark.parametrize("conv_mode", ["fp32", "tf32"])
@pytest.mark.parametrize("gain_tag,gain", [("g1", 1.0), ("g07", 0.7)])
def test_full_model_matches_reference_golden(gain_tag, gain, conv_mode):
    """MonoRecModel.forward through the CUDA path vs the unmodified reference (tests/golden/model_synth_small.npz).

    north-star gate: |delta inverse depth| < 1e-3; additionally every head and the mask are gated relative to their range.
    """
    from monorec_b200.synthetic import make_inputs, to_device
    from tests.helpers import GOLDEN
    from monorec_b200 import conv as C
    g = np.load(GOLDEN / "model_synth_small.npz")
    B, nF, D, H, W, seed, wseed = [int(v) for v in g["cfg"]]
    model, _ = _model_and_sd(gain, wseed)
    old = C.MODE
    C.set_mode(conv_mode)
    try:
        out = model(to_device(make_inputs(B, nF, H, W, seed=seed), DEV))
        torch.cuda.synchronize()
    finally:
        C.set_mode(old)
    dm = np.abs(out["cv_mask"].cpu().numpy() - g[f"{gain_tag}_cv_mask"]).max()
    dd = [np.abs(p.cpu().numpy() - g[f"{gain_tag}_depth{i}"]).max() for i, p in enumerate(out["predicted_inverse_depths"])]
    print(conv_mode, gain_tag, "mask max|d|", dm, "depth max|d|", dd)
    # fp32 arithmetic: gated at 1e-4, a tenth of the north-star 1e-3 (measured on B200: 6.7e-6 for g1, 3.4e-7 for g07; the
    # reference's own fp32-vs-fp64 deviation on this configuration is 1.7e-6 / 1.8e-7, tests/golden/model_fp64.npz).
    # TF32 arithmetic (10-bit mantissa products in every dense layer) is a documented reduced-precision mode, not the parity
    # path: the responsive weight set g07 stays below the north-star 1e-3 (measured 2.0e-4); the g1 weights amplify the
    # product rounding (measured 4.4e-3) and are gated at 1e-2.
    n_res, n_mask = _ref_noise("synth", gain_tag)
    if conv_mode == "fp32":
        tol = max(1e-4, 4 * max(n_res, n_mask))
    else:
        tol = 1e-3 if gain_tag == "g07" else 1e-2
    assert dm < tol and max(dd) < tol
    assert out["result"].shape == (B, 1, H, W) and out["mask"] is out["cv_mask"]
    assert set(["cost_volume", "single_frame_cvs", "image_features", "cv_mask", "predicted_inverse_depths", "result",
                "mask", "inv_depth_min", "inv_depth_max", "cv_depth_steps", "cv_module_time"]) <= set(out.keys())


def test_modules_match_oracle_per_stage(fp32_mode):
    """MaskModule / DepthModule alone (the trainer calls them directly, trainer/monorec_trainer.py:46-89) vs the oracle."""
    from monorec_b200.synthetic import make_inputs, to_device
    from oracle import convnet_oracle as CO
    from oracle import cost_volume_oracle as O
    model, sd = _model_and_sd(0.8, seed=11)
    data = make_inputs(2, 3, 96, 160, seed=9)
    cv, sf = O.cost_volume_torch(data)
    feats = CO.resnet_features(sd, data["keyframe"] + 0.5)
    ref_mask = CO.mask_module(sd, sf, feats)
    ref_depth = CO.depth_module(sd, (1 - ref_mask) * cv, data["keyframe"], feats)
    d = to_device(data, DEV)
    d["single_frame_cvs"] = [s.to(DEV) for s in sf]
    d["image_features"] = [f.to(DEV) for f in feats]
    d = model.att_module(d)
    assert _rel(d["cv_mask"].cpu(), ref_mask) < TOL
    d["cost_volume"] = ((1 - ref_mask) * cv).to(DEV)
    d = model.depth_module(d)
    for p, r in zip(d["predicted_inverse_depths"], ref_depth):
        assert p.shape == r.shape and _rel(p.cpu(), r) < 5 * TOL


# ---------------------------------------------------------------------------------------------------------------------
# tensor-core path (tcgen05 kind::tf32): same layers through PackedConv in "tf32" mode
# ---------------------------------------------------------------------------------------------------------------------
TOL_TF32 = 3e-3   # TF32 products (10-bit mantissa, fp32 accumulate) vs fp32 reference, relative to max|ref| per layer


@pytest.fixture
def tf32_mode():
    from monorec_b200 import conv as C
    old = C.MODE
    C.set_mode("tf32")
    yield
    C.set_mode(old)


@pytest.fixture
def fp32_mode():
    from monorec_b200 import conv as C
    old = C.MODE
    C.set_mode("fp32")
    yield
    C.set_mode(old)


@pytest.mark.parametrize("cin,cout,kh,kw,sy,sx,H,W", [
    (32, 32, 3, 3, 1, 1, 16, 32), (36, 48, 7, 1, 1, 1, 24, 40), (48, 64, 7, 1, 2, 1, 24, 40), (64, 64, 1, 7, 1, 2, 12, 40),
    (64, 128, 5, 1, 2, 1, 20, 24), (128, 128, 1, 5, 1, 2, 10, 24), (192, 256, 3, 1, 2, 1, 18, 16), (32, 24, 3, 3, 1, 1, 9, 21),
    (96, 96, 3, 3, 1, 1, 7, 13), (256, 256, 1, 3, 1, 1, 16, 32), (48, 48, 3, 3, 1, 1, 64, 128),
    (48, 48, 1, 7, 1, 1, 24, 40), (96, 32, 3, 1, 1, 1, 24, 40), (32, 32, 1, 3, 1, 1, 24, 40), (24, 32, 2, 2, 1, 1, 20, 28)])   # halo rows of 14 / 8 / 10 / 9 px
def test_tc_conv_matches_torch(tf32_mode, cin, cout, kh, kw, sy, sx, H, W):
    from monorec_b200 import conv as C
    from oracle.convnet_oracle import conv_same
    g = torch.Generator().manual_seed(cin * 131 + cout + kh)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.leaky_relu(conv_same(x, w, b, (sy, sx)), 0.1)
    layer = C.PackedConv(w.to(DEV), b.to(DEV), (cin,), stride=(sy, sx), act=C.ACT_LEAKY, act_a=0.1)
    assert layer.tc_ok
    out = layer([_nhwc(x).to(DEV)])
    torch.cuda.synchronize()
    assert out.shape == _nhwc(ref).shape
    assert _rel(_nchw(out.cpu()), ref) < TOL_TF32


def test_tc_concat_upconv_refine(tf32_mode):
    from monorec_b200 import conv as C
    from oracle import convnet_oracle as CO
    g = torch.Generator().manual_seed(3)
    a, b_, c = torch.randn(2, 96, 8, 16, generator=g), torch.randn(2, 128, 8, 16, generator=g), torch.randn(2, 36, 8, 16, generator=g)
    cat = torch.cat([a, b_, c], 1)
    srcs = [_nhwc(t).to(DEV) for t in (a, b_, c)]
    conv = torch.nn.Conv2d(260, 64, 3)
    ref = CO.lrelu(CO.conv_same(cat, conv.weight.detach(), conv.bias.detach()))
    out = C.PackedConv(conv.weight.to(DEV), conv.bias.to(DEV), (96, 128, 36), act=C.ACT_LEAKY, act_a=0.1)(srcs)
    assert _rel(_nchw(out.cpu()), ref) < TOL_TF32
    up = torch.nn.Conv2d(260, 96, 2)
    sd = {"u.conv.weight": up.weight.detach(), "u.conv.bias": up.bias.detach()}
    out = C.upconv_layer(up.to(DEV), (96, 128, 36))(srcs)
    assert _rel(_nchw(out.cpu()), CO.upconv(sd, "u", cat)) < TOL_TF32
    ct = torch.nn.ConvTranspose2d(260, 48, 4, stride=2)
    sd = {"r.conv2d_t.weight": ct.weight.detach(), "r.conv2d_t.bias": ct.bias.detach()}
    ref = CO.refine(sd, "r", cat)
    out = C.refine_layer(ct.to(DEV), (96, 128, 36))(srcs)
    assert _rel(_nchw(out.cpu()), ref) < TOL_TF32


def test_fp32_upconv_refine_layers(fp32_mode):
    """the sub-pixel formulations on the CUDA-core kernel: accumulation-order noise only"""
    from monorec_b200 import conv as C
    from oracle import convnet_oracle as CO
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 64, 8, 16, generator=g)
    up = torch.nn.Conv2d(64, 96, 2)
    out = C.upconv_layer(up.to(DEV), (64,))([_nhwc(x).to(DEV)])
    assert _rel(_nchw(out.cpu()), CO.upconv({"u.conv.weight": up.weight.detach().cpu(), "u.conv.bias": up.bias.detach().cpu()}, "u", x)) < TOL
    ct = torch.nn.ConvTranspose2d(64, 48, 4, stride=2)
    ref = CO.refine({"r.conv2d_t.weight": ct.weight.detach(), "r.conv2d_t.bias": ct.bias.detach()}, "r", x)
    out = C.refine_layer(ct.to(DEV), (64,))([_nhwc(x).to(DEV)])
    assert _rel(_nchw(out.cpu()), ref) < TOL


def test_modules_tf32_vs_oracle(tf32_mode):
    """Mask / depth stacks on the tensor cores vs the fp32 oracle on identical inputs (stated TF32 tolerance)."""
    from monorec_b200.synthetic import make_inputs, to_device
    from oracle import convnet_oracle as CO
    from oracle import cost_volume_oracle as O
    model, sd = _model_and_sd(0.8, seed=11)
    data = make_inputs(1, 2, 96, 160, seed=9)
    cv, sf = O.cost_volume_torch(data)
    feats = CO.resnet_features(sd, data["keyframe"] + 0.5)
    ref_mask = CO.mask_module(sd, sf, feats)
    ref_depth = CO.depth_module(sd, (1 - ref_mask) * cv, data["keyframe"], feats)
    d = to_device(data, DEV)
    d["single_frame_cvs"] = [s.to(DEV) for s in sf]
    d["image_features"] = [f.to(DEV) for f in feats]
    d = model.att_module(d)
    torch.cuda.synchronize()
    em = _rel(d["cv_mask"].cpu(), ref_mask)
    d["cost_volume"] = ((1 - ref_mask) * cv).to(DEV)
    d = model.depth_module(d)
    torch.cuda.synchronize()
    ed = [_rel(p.cpu(), r) for p, r in zip(d["predicted_inverse_depths"], ref_depth)]
    print("tf32 stacks: mask rel err", em, "depth rel err", ed)
    assert em < 1e-2 and max(ed) < 1e-2


def test_tc_halo_variant_in_subprocess():
    """The opt-in halo-reuse kernel (MONOREC_B200_TC_HALO=1; one input box per tile, resident weights) computes the same
    layers; the switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, MONOREC_B200_TC_HALO="1")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-m", "gpu", "-k", "tc_conv_matches_torch or tc_concat"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]


def test_full_size_model_modes_and_graph_replay():
    """BASELINE config 3 shape (256x512, D=32, F=4): tensor-core vs CUDA-core arithmetic agree within the TF32 tolerance,
    and the CUDA-graph replay reproduces the eager forward bit for bit."""
    from monorec_b200 import conv as C
    from monorec_b200.model import GraphedMonoRec
    from monorec_b200.synthetic import make_inputs, to_device
    model, _ = _model_and_sd(0.7)
    data = to_device(make_inputs(2, 4, 256, 512, seed=23), DEV)
    old = C.MODE
    try:
        C.set_mode("fp32")
        ref = model(dict(data))
        ref_res, ref_mask = ref["result"].clone(), ref["cv_mask"].clone()
        C.set_mode("tf32")
        out = model(dict(data))
        res, mask = out["result"].clone(), out["cv_mask"].clone()
        g = GraphedMonoRec(model, data)
        rep = g(data)
        torch.cuda.synchronize()
        assert torch.equal(rep["result"], res) and torch.equal(rep["cv_mask"], mask)
    finally:
        C.set_mode(old)
    dr, dm = (res - ref_res).abs().max().item(), (mask - ref_mask).abs().max().item()
    print("full size tf32 vs fp32: inverse depth max|d|", dr, "mask max|d|", dm)
    assert dr < 1e-3 and dm < 5e-3
    assert res.shape == (2, 1, 256, 512) and [p.shape[-1] for p in out["predicted_inverse_depths"]] == [512, 256, 128, 64]


# ---------------------------------------------------------------------------------------------------------------------
# half-precision storage (BASELINE config 3): tcgen05 kind::f16, half NHWC activations and weights, fp32 accumulation
# ---------------------------------------------------------------------------------------------------------------------
TOL_F16 = 4e-3   # half inputs/weights/outputs (10-bit mantissa each) vs fp32 reference, relative to max|ref| per layer


@pytest.fixture
def f16_mode():
    from monorec_b200 import conv as C
    old = C.MODE
    C.set_mode("f16")
    yield
    C.set_mode(old)


@pytest.mark.parametrize("cin,cout,kh,kw,sy,sx,H,W", [
    (32, 32, 3, 3, 1, 1, 16, 32), (40, 48, 7, 1, 1, 1, 24, 40), (48, 64, 7, 1, 2, 1, 24, 40), (64, 64, 1, 7, 1, 2, 12, 40),
    (128, 128, 1, 5, 1, 2, 10, 24), (192, 256, 3, 1, 2, 1, 18, 16), (32, 24, 3, 3, 1, 1, 9, 21), (96, 96, 3, 3, 1, 1, 7, 13),
    (48, 48, 3, 3, 1, 1, 64, 128), (48, 48, 1, 7, 1, 1, 24, 40), (96, 32, 3, 1, 1, 1, 24, 40), (32, 32, 1, 3, 1, 1, 24, 40)])
def test_f16_conv_matches_torch(f16_mode, cin, cout, kh, kw, sy, sx, H, W):
    from monorec_b200 import conv as C
    from oracle.convnet_oracle import conv_same
    g = torch.Generator().manual_seed(cin * 17 + cout + kw)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, kh, kw, generator=g) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.leaky_relu(conv_same(x, w, b, (sy, sx)), 0.1)
    layer = C.PackedConv(w.to(DEV), b.to(DEV), (cin,), stride=(sy, sx), act=C.ACT_LEAKY, act_a=0.1)
    assert layer.tc_ok_f16
    out = layer([_nhwc(x).to(DEV).half()])
    torch.cuda.synchronize()
    assert out.dtype == torch.float16 and out.shape == _nhwc(ref).shape
    assert _rel(_nchw(out.float().cpu()), ref) < TOL_F16


def test_f16_helpers_and_subpixel(f16_mode):
    from monorec_b200 import conv as C
    from oracle import convnet_oracle as CO
    g = torch.Generator().manual_seed(8)
    x = torch.randn(4, 32, 12, 20, generator=g)
    xh = C.nchw_to_nhwc(x.to(DEV), dtype=torch.float16)
    assert xh.dtype == torch.float16 and torch.equal(xh.cpu(), _nhwc(x).half())
    assert torch.equal(C.maxpool2(xh).cpu(), _nhwc(F.max_pool2d(x.half().float(), 2)).half())
    assert torch.equal(C.max_over_frames(xh, 2).cpu(), torch.maximum(_nhwc(x).half()[:2], _nhwc(x).half()[2:]))
    cl = x.to(DEV).contiguous(memory_format=torch.channels_last)
    assert torch.equal(C.nchw_to_nhwc(cl, dtype=torch.float16).cpu(), _nhwc(x).half())
    # tiled fast path (C % 32 == 0, H*W % 128 == 0): plain, and into a channel slice with the (1 - mask) product
    y = torch.randn(3, 64, 16, 24, generator=g)
    m = torch.rand(3, 1, 16, 24, generator=g)
    assert torch.equal(C.nchw_to_nhwc(y.to(DEV), dtype=torch.float16).cpu(), _nhwc(y).half())
    buf = torch.zeros(3, 16, 24, 72, device=DEV, dtype=torch.float16)
    C.nchw_to_nhwc(y.to(DEV), out=buf, out_coff=8, one_minus=m.to(DEV))
    assert torch.equal(buf[..., 8:].cpu(), _nhwc(y * (1.0 - m)).half()) and float(buf[..., :8].abs().max()) == 0.0
    ct = torch.nn.ConvTranspose2d(32, 48, 4, stride=2)
    ref = CO.refine({"r.conv2d_t.weight": ct.weight.detach(), "r.conv2d_t.bias": ct.bias.detach()}, "r", x)
    out = C.refine_layer(ct.to(DEV), (32,))([xh])
    assert out.dtype == torch.float16 and _rel(_nchw(out.float().cpu()), ref) < TOL_F16
    head = torch.nn.Conv2d(32, 1, 3)
    refh = torch.abs(torch.tanh(CO.conv_same(x, head.weight.detach(), head.bias.detach())))
    outh = C.PackedConv(head.weight.to(DEV), head.bias.to(DEV), (32,), act=C.ACT_ABSTANH, act_a=0.0, act_b=1.0, allow_tc=False)([xh], final=True)
    assert outh.dtype == torch.float32 and _rel(_nchw(outh.cpu()), refh) < TOL_F16


def test_full_model_f16_matches_reference_golden(f16_mode):
    """BASELINE config 3 arithmetic on the golden model: |delta inverse depth| < 1e-3 against the fp32 reference."""
    from monorec_b200.synthetic import make_inputs, to_device
    from tests.helpers import GOLDEN
    g = np.load(GOLDEN / "model_synth_small.npz")
    B, nF, D, H, W, seed, wseed = [int(v) for v in g["cfg"]]
    model, _ = _model_and_sd(0.7, wseed)
    out = model(to_device(make_inputs(B, nF, H, W, seed=seed), DEV))
    torch.cuda.synchronize()
    dm = np.abs(out["cv_mask"].cpu().numpy() - g["g07_cv_mask"]).max()
    dd = [np.abs(p.cpu().numpy() - g[f"g07_depth{i}"]).max() for i, p in enumerate(out["predicted_inverse_depths"])]
    print("f16 g07 mask max|d|", dm, "depth max|d|", dd)
    assert dm < 2e-3 and max(dd) < 1e-3


def _ref_noise(cfg, gain_tag):
    """The reference's own fp32-vs-fp64 deviation on (result, cv_mask) for this configuration (tests/golden/model_fp64.npz,
    written by `make_golden.py --only-model-fp64`): no fp32 implementation can be held closer to the fp32 reference than that."""
    from tests.helpers import GOLDEN
    n = np.load(GOLDEN / "model_fp64.npz")[f"{cfg}_{gain_tag}_noise"]
    return float(n[0]), float(n[1])


@pytest.mark.parametrize("mode", ["fp32", "tf32", "f16"])
@pytest.mark.parametrize("gain_tag,gain", [("g1", 1.0), ("g07", 0.7)])
def test_full_model_on_bundled_sample(mode, gain_tag, gain):
    """The north-star parity sentence on the CUDA path: full MonoRecModel on the bundled KITTI sample (256x512, 2 source frames)
    against the unmodified reference (tests/golden/model_kitti_sample.npz, seeded weights).

    fp32 arithmetic (the parity path): |delta inverse depth| < max(1e-4, 4 x the reference's own fp32-vs-fp64 deviation), i.e.
    ten times tighter than the north-star 1e-3 for the g1 weights (measured on B200: 5.5e-6; reference noise 2.3e-6).  The g07
    weight set is chaotic on this image -- the reference itself moves by 8.3e-3 between fp32 and fp64 (tests/golden/
    model_fp64.npz) -- so its max-norm gate is 4 x that and the informative gate is the share of pixels within 1e-3 (measured
    0.99924 in every mode; max|d| 4.3e-3, inside the reference's own noise).
    tf32 / f16 are reduced-precision arithmetic modes (10-bit mantissa products): gated at 1e-2 (measured 4.2e-3 / 3.9e-3 for
    g1) and on the pixel share within 1e-3 (measured 0.961 / 0.959); cv_mask at 2e-2 (measured 7.6e-3 / 1.08e-2)."""
    from monorec_b200 import conv as K
    from monorec_b200.model import MonoRecModel
    from monorec_b200.synthetic import seeded_state_dict, to_device
    from tests.helpers import GOLDEN, kitti_sample_dict
    g = np.load(GOLDEN / "model_kitti_sample.npz")
    data, _ = kitti_sample_dict()
    old = K.MODE
    K.set_mode(mode)
    try:
        model = MonoRecModel()
        model.load_state_dict(seeded_state_dict(model, seed=int(g["wseed"][0]), gain=gain))
        model = model.to(DEV).eval()
        with torch.no_grad():
            out = model(to_device(data, DEV))
        torch.cuda.synchronize()
    finally:
        K.set_mode(old)
    dr = np.abs(out["result"].float().cpu().numpy() - g[f"{gain_tag}_result"])
    dmask = np.abs(out["cv_mask"].float().cpu().numpy() - g[f"{gain_tag}_cv_mask"].astype(np.float32))
    n_res, n_mask = _ref_noise("kitti", gain_tag)
    share = float((dr < 1e-3).mean())
    print(f"bundled sample {mode} {gain_tag}: result max|d| {dr.max():.3e} (reference fp32-vs-fp64 {n_res:.1e}), "
          f"share within 1e-3: {share:.5f}, mask max|d| {dmask.max():.3e} (reference {n_mask:.1e})")
    if mode == "fp32":
        assert dr.max() < max(1e-4, 4 * n_res) and share > 0.999
        assert dmask.max() < max(1e-3, 4 * n_mask)      # the golden mask is stored as half (5e-4 quantisation)
    else:
        assert dr.max() < max(1e-2, 4 * n_res) and share > (0.99 if gain_tag == "g07" else 0.9)
        # the mask is a sigmoid over half / TF32 activations: its worst pixel moved between 8.4e-3 and 1.08e-2 when only the
        # accumulation order of some layers changed (tap-major -> chunk-major), so the max-norm gate is 2e-2 and the informative
        # gate is the share of pixels within 5e-3
        mshare = float((dmask < 5e-3).mean())
        print(f"    mask share within 5e-3: {mshare:.5f}")
        assert dmask.max() < max(2e-2, 4 * n_mask) and mshare > 0.99


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_stem_maxpool_matches_torch(dtype):
    """MaxPool2d(3, stride 2, padding 1) on channels-last tensors (the ResNet stem pool), bit for bit, odd sizes included."""
    from monorec_b200 import conv as C
    g = torch.Generator().manual_seed(11)
    for (B, Cc, H, W) in [(2, 64, 32, 64), (1, 16, 17, 23), (3, 8, 5, 2)]:
        x = torch.randn(B, Cc, H, W, generator=g).to(DEV, dtype).contiguous(memory_format=torch.channels_last)
        out = C.maxpool3s2_channels_last(x)
        ref = F.max_pool2d(x.float(), 3, 2, 1).to(dtype)
        assert out.shape == ref.shape and torch.equal(out, ref)
        assert out.is_contiguous(memory_format=torch.channels_last) or out.shape[2] * out.shape[3] == 1 or Cc == 1
