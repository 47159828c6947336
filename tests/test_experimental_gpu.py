"""GPU tests of the opt-in variants that were written without GPU minutes left (residual epilogue, 3x3/s2 max-pool, ResNet
trunk on the conv engine).  They run only with MONOREC_B200_EXPERIMENTAL=1 so that the regular `-m gpu` suite keeps covering
exactly the measured default path; tools/gpu_next_round.sh sets the switches:

    MONOREC_B200_EXPERIMENTAL=1 MONOREC_B200_TC_EPI=1 MONOREC_B200_TRUNK=engine python -m pytest tests/test_experimental_gpu.py -m gpu
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MONOREC_B200_EXPERIMENTAL") != "1", reason="opt-in: MONOREC_B200_EXPERIMENTAL=1")]
DEV = "cuda:0"
STAGED = os.environ.get("MONOREC_B200_TC_EPI") == "1"


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


@pytest.mark.parametrize("H,W,C", [(16, 24, 64), (15, 21, 8), (2, 2, 16)])
def test_maxpool3s2_matches_torch(H, W, C):
    from monorec_b200 import conv as K
    x = torch.randn(3, C, H, W, generator=torch.Generator().manual_seed(H * W + C))
    ref = F.max_pool2d(x, 3, 2, 1)
    assert torch.equal(_nchw(K.maxpool3s2(_nhwc(x).to(DEV)).cpu()), ref)
    xh = x.half()
    assert torch.equal(_nchw(K.maxpool3s2(_nhwc(xh).to(DEV)).cpu()), F.max_pool2d(xh.float(), 3, 2, 1).half())


@pytest.mark.skipif(not STAGED, reason="the residual input lives in the staged epilogue (MONOREC_B200_TC_EPI=1)")
@pytest.mark.parametrize("mode,tol", [("tf32", 3e-3), ("f16", 4e-3)])
@pytest.mark.parametrize("cin,cout,stride,H,W", [(64, 64, 1, 16, 32), (64, 128, 2, 18, 20), (256, 512, 2, 8, 16)])
def test_residual_conv_matches_torch(mode, tol, cin, cout, stride, H, W):
    """relu(conv3x3(x) + bias + residual) with PyTorch padding 1 -- the second conv of a BasicBlock; 512 channels = 2 slices."""
    from monorec_b200 import conv as K
    old = K.MODE
    K.set_mode(mode)
    try:
        g = torch.Generator().manual_seed(cin + cout)
        x = torch.randn(2, cin, H, W, generator=g)
        w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
        b = torch.randn(cout, generator=g)
        ref_lin = F.conv2d(x, w, b, stride=stride, padding=1)
        res = torch.randn(ref_lin.shape, generator=g)
        ref = F.relu(ref_lin + res)
        dt = K.act_dtype()
        xs = _nhwc(x).to(DEV, dt)
        rs = _nhwc(res).to(DEV, dt)
        out = torch.empty(rs.shape, device=DEV, dtype=dt)
        for c0 in range(0, cout, 256):
            layer = K.PackedConv(w[c0:c0 + 256].to(DEV), b[c0:c0 + 256].to(DEV), (cin,), stride=(stride, stride), pad=(1, 1),
                                 act=K.ACT_LEAKY, act_a=0.0)
            layer([xs], out=out, out_coff=c0, residual=rs)
        torch.cuda.synchronize()
        assert _rel(_nchw(out.float().cpu()), ref) < tol
    finally:
        K.set_mode(old)


@pytest.mark.skipif(not STAGED, reason="needs MONOREC_B200_TC_EPI=1")
@pytest.mark.parametrize("mode,tol", [("tf32", 1e-2), ("f16", 1.5e-2)])
def test_engine_trunk_matches_cudnn_trunk(mode, tol):
    """ResnetEncoder on the conv engine against the same (folded) trunk on cuDNN fp32, all five feature maps."""
    from monorec_b200 import conv as K
    from monorec_b200.model import ResnetEncoder
    old = K.MODE
    K.set_mode(mode)
    try:
        torch.manual_seed(3)
        enc = ResnetEncoder(18, pretrained=False).eval()
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.3)
                m.running_var.uniform_(0.5, 2.0)
                m.bias.data.normal_(0, 0.2)
        enc = enc.to(DEV)
        x = torch.rand(2, 3, 64, 128, device=DEV)
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            ref = [t.float().clone() for t in enc._forward_folded(x)]
            got = enc._forward_engine(x)
        torch.cuda.synchronize()
        for a, b in zip(ref, got):
            assert a.shape == b.shape and _rel(b.float().cpu(), a.cpu()) < tol
    finally:
        K.set_mode(old)


def test_cost_volume_golden_d64_f6():
    """CUDA cost volume vs the reference's own output for 64 planes x 6 source frames (tests/golden/cv_synth_d64f6.npz); the
    same gates as tests/test_cost_volume_gpu.py::test_golden_small_full_tensors -- moves there once it has run on a GPU."""
    from tests.helpers import compare_volumes, synth_small_dict
    from tests.test_cost_volume_gpu import _run
    data, D, ref_cv, ref_sf = synth_small_dict("d")
    cv, sf = _run(data, steps=D)
    print(compare_volumes(cv, sf, ref_cv, ref_sf))


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-3), ("tf32", 1e-3), ("f16", 1e-3)])
@pytest.mark.parametrize("gain_tag,gain", [("g1", 1.0), ("g07", 0.7)])
def test_full_model_on_bundled_sample(mode, tol, gain_tag, gain):
    """|delta inverse depth| < 1e-3 against the unmodified reference on the bundled KITTI sample (tests/golden/model_kitti_sample.npz,
    seeded weights): the north-star parity sentence on the CUDA path.  Moves to the regular suite once it has run on a GPU."""
    import numpy as np
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
        dres = np.abs(out["result"].float().cpu().numpy() - g[f"{gain_tag}_result"]).max()
        dmask = np.abs(out["cv_mask"].float().cpu().numpy() - g[f"{gain_tag}_cv_mask"].astype(np.float32)).max()
        print(mode, gain_tag, "result", dres, "mask", dmask)
        assert dres < tol and dmask < 5e-3
    finally:
        K.set_mode(old)
