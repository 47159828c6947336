"""the photometric reprojection loss (model/loss_functions/common_losses.py:16-114), forward and
backward.

CPU: the torch oracle against tests/golden/reprojection.npz (errors and autograd gradients of the UNMODIFIED reference), the
hand-derived closed form against autograd in float64, the host-side argument checks.  GPU: mr_reprojection_loss_fwd / _bwd
through the C ABI against the golden and against the oracle at other sizes.

Gates.  The reference's own fp32 rounding noise (its fp32 run against its float64 run, measured with the oracle, which
reproduces the reference bit for bit): on the golden inputs errors 1.2e-5, gradient 1.4e-3 absolute at max |gradient| 41; at
256x512 errors 7.4e-5, two winner flips, and gradient differences up to 17 (max |gradient| 869) on 26 pixels.  Those large
differences sit at the loss's non-differentiable points -- a window at a bound of SSIM's clamp (layers.py:139), x == y of the
L1 term, a sample on a pixel boundary -- where two correct fp32 roundings take different branches.  The CUDA path is another
fp32 rounding of the same function (one homography instead of back-projection + normalised grid), so:
  * errors agree to 3e-4 (north-star tolerance 1e-3) wherever the winning frame agrees,
  * the discrete decisions (mask, auto-mask, winning frame) flip on at most 0.1 % of the pixels (none flips on the golden inputs),
  * gradients agree to 1e-2 + 5e-4 max |gradient| at every pixel that is not within reach of such a point; the float64 closed form
    marks them (kink_eps 1e-4: 0.1-0.4 % of the pixels, at most 1 %).
"""
import warnings

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, reprojection_inputs

CASES = {"plain": dict(use_mono=True, use_stereo=False, automasking=False),
         "auto": dict(use_mono=True, use_stereo=True, automasking=True),
         "stereo_border": dict(use_mono=False, use_stereo=True, automasking=False, border=3)}
ERR_TOL, GRAD_ABS, GRAD_REL, FLIP_SHARE, KINK_EPS, KINK_SHARE = 3e-4, 1e-2, 5e-4, 1e-3, 1e-4, 1e-2


def _oracle(invd, data, wts, dtype=torch.float32, **kw):
    from oracle import reprojection_oracle as RO
    d = {k: ([t.to(dtype) for t in v] if isinstance(v, list) else v.to(dtype)) for k, v in data.items()}
    p = invd.to(dtype).clone().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e, idx = RO.reprojection_errors(p, d, **kw)
        inf = torch.isinf(e)
        (torch.where(inf, torch.zeros_like(e), e) * wts.to(dtype)).sum().backward()
    return e.detach(), idx, p.grad


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_matches_reference_golden(tag):
    g = np.load(GOLDEN / "reprojection.npz")
    data, invd, wts = reprojection_inputs()
    assert np.array_equal(g["invd"], invd.numpy()) and np.array_equal(g["weights"], wts.numpy())   # same seeded inputs
    e, idx, grad = _oracle(invd, data, wts, **CASES[tag])
    ref_e, ref_g = torch.from_numpy(g[f"errors_{tag}"]), torch.from_numpy(g[f"grad_{tag}"])
    assert torch.equal(torch.isinf(e), torch.isinf(ref_e))
    fin = ~torch.isinf(e)
    assert torch.allclose(e[fin], ref_e[fin], rtol=0, atol=1e-6)
    assert torch.allclose(grad, ref_g, rtol=1e-5, atol=1e-4)
    assert ((idx == -1) == torch.isinf(ref_e)).all()
    from oracle import reprojection_oracle as RO
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        red = float(RO.reprojection_loss(invd, data, reduce=True, **CASES[tag]))
    assert abs(red - float(g[f"reduced_{tag}"])) < 1e-6
    if tag != "plain":
        assert 0.05 < float(torch.isinf(ref_e).float().mean()) < 0.5          # the masks are exercised


@pytest.mark.parametrize("tag", list(CASES))
def test_closed_form_gradient_matches_autograd_fp64(tag):
    """The hand-derived backward (the formulation csrc/reprojection.cu implements) against torch autograd of the reference's
    primitives, both in float64."""
    from oracle import reprojection_oracle as RO
    data, invd, wts = reprojection_inputs()
    e, idx, grad = _oracle(invd, data, wts, dtype=torch.float64, **CASES[tag])
    be, wi, gr = RO.reprojection_closed_form(invd, data, grad_errors=wts.numpy(), **CASES[tag])
    assert np.array_equal(wi, idx.numpy())
    fin = wi >= 0
    assert np.abs(be[fin] - e.numpy()[fin]).max() < 1e-10
    assert np.isinf(be[~fin]).all()
    assert np.abs(gr - grad.numpy()).max() < 1e-8


def test_unsupported_arguments_raise():
    from monorec_b200 import losses as L
    data, invd, _ = reprojection_inputs()
    with pytest.raises(NotImplementedError):
        L.reprojection_loss(invd, data, combine_frames="avg")
    with pytest.raises(NotImplementedError):
        L.reprojection_loss(invd, data, mono_auto=True)
    with pytest.raises(NotImplementedError):
        L.reprojection_loss(invd, data, error_function=lambda a, b, m=None: a)
    with pytest.raises(Exception, match="CUDA"):
        L.reprojection_loss(invd, data)                                      # CPU tensors: no fallback


def _cuda_run(invd, data, wts, reduce=False, **kw):
    from monorec_b200 import losses as L
    from monorec_b200.synthetic import to_device
    dev = "cuda:0"
    d = to_device(data, dev)
    p = invd.to(dev).clone().requires_grad_(True)
    if reduce:
        loss = L.reprojection_loss(p, d, error_function=L.compute_errors, reduce=True, **kw)
        loss.backward()
        return loss.detach().cpu(), None, p.grad.cpu()
    e, win = L.reprojection_errors(p, d, kw.get("automasking", False), kw.get("use_mono", True), kw.get("use_stereo", False),
                                   kw.get("border", 0))
    inf = torch.isinf(e)
    (torch.where(inf, torch.zeros_like(e), e) * wts.to(dev)).sum().backward()
    return e.detach().cpu(), win.cpu(), p.grad.cpu()


def _kinks(invd, data, wts, **kw):
    from oracle import reprojection_oracle as RO
    _, _, _, kink = RO.reprojection_closed_form(invd, data, grad_errors=wts.numpy(), kink_eps=KINK_EPS, **kw)
    return torch.from_numpy(kink)


def _compare(e, win, grad, ref_e, ref_idx, ref_grad, kink, tag):
    assert float(kink.float().mean()) <= KINK_SHARE
    flips = (win.long() != ref_idx.long())
    share = float(flips.float().mean())
    assert share <= FLIP_SHARE, f"{tag}: {share:.4%} of the pixels pick another frame / mask"
    same = ~flips
    fin = same & (ref_idx >= 0)
    assert torch.isinf(e[same & (ref_idx < 0)]).all()
    err = float((e[fin] - ref_e[fin]).abs().max())
    assert err <= ERR_TOL, f"{tag}: errors differ by {err:.3g}"
    # gradients: a flipped pixel changes the gradient of its 3x3 neighbourhood, leave those and the marked kinks out
    near = (torch.nn.functional.max_pool2d(flips.float().unsqueeze(1), 3, 1, 1) > 0) | kink.unsqueeze(1)
    dg = float(((grad - ref_grad).abs() * (~near)).max())
    gmax = float(ref_grad.abs().max())
    assert dg <= GRAD_ABS + GRAD_REL * gmax, f"{tag}: gradients differ by {dg:.3g} (max |gradient| {gmax:.3g})"
    return share, err, dg


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_cuda_matches_reference_golden(tag):
    g = np.load(GOLDEN / "reprojection.npz")
    data, invd, wts = reprojection_inputs()
    e, win, grad = _cuda_run(invd, data, wts, **CASES[tag])
    ref_e, ref_g = torch.from_numpy(g[f"errors_{tag}"]), torch.from_numpy(g[f"grad_{tag}"])
    _, ref_idx, _ = _oracle(invd, data, wts, **CASES[tag])                  # the winner map (the reference does not return it)
    share, err, dg = _compare(e, win, grad, ref_e, ref_idx, ref_g, _kinks(invd, data, wts, **CASES[tag]), tag)
    print(f"reprojection golden {tag}: flips {share:.5f}, max |d errors| {err:.3g}, max |d grad| {dg:.3g}")
    # reduce=True: mask_mean of the finite errors (common_losses.py:110-111) and its gradient
    red, _, rgrad = _cuda_run(invd, data, wts, reduce=True, **CASES[tag])
    assert abs(float(red) - float(g[f"reduced_{tag}"])) < 1e-5
    n_fin = float((~torch.isinf(ref_e)).sum())
    _, _, og = _oracle(invd, data, torch.full_like(wts, 1.0 / n_fin), **CASES[tag])
    kink = _kinks(invd, data, wts, **CASES[tag]).unsqueeze(1)
    assert float(((rgrad - og).abs() * ~kink).max()) <= (GRAD_ABS + GRAD_REL * float(og.abs().max()) * n_fin) / n_fin


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(1, 3, 50, 70, 5, dict(automasking=True)), (2, 1, 33, 95, 6, dict(border=2)),
                                 (1, 4, 256, 512, 100, dict(automasking=True))])
def test_cuda_matches_oracle_other_sizes(cfg):
    """Ragged sizes (tiles that straddle the image edge), one and four source frames, BASELINE config 2's image size."""
    from monorec_b200.synthetic import make_inputs
    B, Fn, H, W, seed, kw = cfg
    data = make_inputs(B, Fn, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed)
    yy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1) / H
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W) / W
    invd = (0.15 + 0.12 * torch.sin(4.0 * xx + 3.0 * yy) + 0.01 * (torch.rand(B, 1, H, W, generator=g) - 0.5)).clamp(0.01, 0.3)
    wts = torch.rand(B, H, W, generator=g) + 0.5
    e, win, grad = _cuda_run(invd, data, wts, **kw)
    ref_e, ref_idx, ref_g = _oracle(invd, data, wts, **kw)
    share, err, dg = _compare(e, win, grad, ref_e, ref_idx, ref_g, _kinks(invd, data, wts, **kw), str(cfg))
    print(f"reprojection {cfg}: flips {share:.5f}, max |d errors| {err:.3g}, max |d grad| {dg:.3g}, inf share "
          f"{float(torch.isinf(ref_e).float().mean()):.3f}")
    # pixels without a usable frame get no gradient from their own window; the run is deterministic
    e2, win2, grad2 = _cuda_run(invd, data, wts, **kw)
    assert torch.equal(e, e2) and torch.equal(win, win2) and torch.equal(grad, grad2)


@pytest.mark.gpu
def test_c_abi_argument_checks():
    from monorec_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(1, 3, 8, 8, device="cuda:0")
    proj = torch.zeros(1, 1, 12, device="cuda:0")
    inv = torch.ones(1, 1, 8, 8, device="cuda:0")
    e = torch.empty(1, 8, 8, device="cuda:0")
    w = torch.empty(1, 8, 8, device="cuda:0", dtype=torch.int32)
    rc = lib.mr_reprojection_loss_fwd(t.data_ptr(), _lib.ptr_array([t]), proj.data_ptr(), inv.data_ptr(), 1, 1, 8, 8, 0, 4,
                                      e.data_ptr(), w.data_ptr(), None)
    assert rc != 0 and b"border" in lib.mr_last_error()
    rc = lib.mr_reprojection_loss_fwd(t.data_ptr(), _lib.ptr_array([t]), proj.data_ptr(), inv.data_ptr(), 1, 9, 8, 8, 0, 0,
                                      e.data_ptr(), w.data_ptr(), None)
    assert rc != 0
