"""the fused sparse-metric pass and the uint8 image normalisation.

CPU: the numpy oracle against the reference's own outputs (tests/golden/metrics.npz, written by make_golden.py --only-metrics
from the unmodified model/metric_functions/sparse_metrics.py).  GPU: the CUDA pass through the C ABI against the golden and
against the oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN

NAMES = ("a1", "a2", "a3", "rmse", "rmse_log", "abs_rel", "sq_rel")
CASES = {"plain": dict(), "roi_md": dict(roi=[4, 44, 8, 72], max_distance=80.0),
         "onlyvalid": dict(roi=None, max_distance=50.0, pred_all_valid=False),
         "onlydynamic": dict(roi=None, max_distance=80.0, use_cvmask=True)}


def _oracle(g, kw):
    from oracle import metrics_oracle as MO
    kw = dict(kw)
    mv = g["mvobj"] if kw.pop("use_cvmask", False) else None
    return MO.sparse_metrics(g["pred"], g["gt"], mvobj_mask=mv, **kw)


@pytest.mark.parametrize("tag", list(CASES))
def test_metrics_oracle_matches_reference_golden(tag):
    g = np.load(GOLDEN / "metrics.npz")
    got = _oracle(g, CASES[tag])
    np.testing.assert_allclose([got[n] for n in NAMES], g[f"case_{tag}"], rtol=2e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_cuda_metrics_match_reference_golden(tag):
    from monorec_b200 import metrics as M
    g = np.load(GOLDEN / "metrics.npz")
    d = {"result": torch.from_numpy(g["pred"]).cuda(), "target": torch.from_numpy(g["gt"]).cuda(),
         "mvobj_mask": torch.from_numpy(g["mvobj"]).cuda()}
    out = M.sparse_metrics(d, **CASES[tag]).cpu().numpy()
    np.testing.assert_allclose(out, g[f"case_{tag}"], rtol=5e-6, atol=1e-7)
    # the reference-named functions share the one fused pass
    assert float(M.abs_rel_sparse_metric(d, **CASES[tag])) == float(out[5])
    assert d["_mr_metrics_cache"][1].data_ptr() == M.sparse_metrics(d, **CASES[tag]).data_ptr()


@pytest.mark.gpu
def test_cuda_metrics_match_oracle_at_full_size():
    from monorec_b200 import metrics as M
    from oracle import metrics_oracle as MO
    gen = torch.Generator().manual_seed(5)
    B, H, W = 4, 256, 512
    pred = torch.rand(B, 1, H, W, generator=gen) * 0.3 + 0.002
    gt = (pred * (1 + 0.2 * torch.randn(B, 1, H, W, generator=gen))).clamp_min(1e-3)
    gt[torch.rand(B, 1, H, W, generator=gen) > 0.05] = 0.0
    d = {"result": pred.cuda(), "target": gt.cuda()}
    out = M.sparse_metrics(d, roi=[40, 250, 20, 500], max_distance=80.0).cpu().numpy()
    ref = MO.sparse_metrics(pred.numpy(), gt.numpy(), roi=[40, 250, 20, 500], max_distance=80.0)
    np.testing.assert_allclose(out, [ref[n] for n in NAMES], rtol=5e-6, atol=1e-7)


@pytest.mark.gpu
def test_images_u8_to_f32_matches_loader_formula():
    from monorec_b200 import metrics as M
    gen = torch.Generator().manual_seed(2)
    u8 = torch.randint(0, 256, (2, 37, 61, 3), generator=gen, dtype=torch.uint8)
    box = (5, 3, 53, 35)                                   # PIL crop box (left, upper, right, lower)
    ref = (u8[:, box[1]:box[3], box[0]:box[2]].to(torch.float32) / 255 - .5).permute(0, 3, 1, 2)
    out = M.images_u8_to_f32(u8.cuda(), crop_box=box).cpu()
    assert torch.equal(out, ref)
    assert torch.equal(M.images_u8_to_f32(u8.cuda()).cpu(), (u8.to(torch.float32) / 255 - .5).permute(0, 3, 1, 2))
