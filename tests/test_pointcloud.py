"""mask dilation, window vote, back-projection and ordered compaction of the point-cloud export.

CPU: the torch oracle against the unmodified reference PLYSaver (tests/golden/pointcloud.npz).  GPU: the CUDA path through
the C ABI against the golden, vertex by vertex and in the reference's order."""
import io

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN

CASES = {"plain": (None, False), "roi_vote": ([4, 36, 6, 60], True)}


def _inputs():
    g = np.load(GOLDEN / "pointcloud.npz")
    t = {k: torch.from_numpy(g[k]) for k in ("inv_depth", "image", "K", "pose", "cv_masks", "keeps")}
    return g, t


@pytest.mark.parametrize("tag", list(CASES))
def test_pointcloud_oracle_matches_reference_golden(tag):
    from oracle import pointcloud_oracle as PO
    g, t = _inputs()
    roi, vote = CASES[tag]
    keeps = [PO.keep_mask(m) for m in t["cv_masks"]]
    assert all(torch.equal(k, r) for k, r in zip(keeps, t["keeps"]))
    v = PO.add_depthmap(t["inv_depth"], t["image"], t["K"], t["pose"], keep_masks=keeps if vote else (), min_d=3, max_d=30, roi=roi)
    ref = torch.from_numpy(g[f"vertices_{tag}"])
    assert v.shape == ref.shape and v.shape[0] > 100
    assert torch.allclose(v, ref, rtol=1e-6, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_cuda_pointcloud_matches_reference_golden(tag):
    from monorec_b200 import pointcloud as PC
    g, t = _inputs()
    roi, vote = CASES[tag]
    dev = "cuda:0"
    keeps = [PC.keep_mask(m.to(dev)) for m in t["cv_masks"]]
    assert all(torch.equal(k.cpu(), r) for k, r in zip(keeps, t["keeps"]))
    B, _, H, W = t["inv_depth"].shape
    saver = PC.PLYSaver(H, W, min_d=3, max_d=30, batch_size=B, roi=roi, dropout=0)
    saver.add_depthmap(t["inv_depth"].to(dev), t["image"].to(dev), t["K"].to(dev), t["pose"].to(dev),
                       keep_masks=keeps if vote else ())
    ref = torch.from_numpy(g[f"vertices_{tag}"])
    v = saver.vertices.cpu()
    assert v.shape == ref.shape
    assert torch.allclose(v, ref, rtol=2e-6, atol=2e-5)          # same vertices in the same order
    # a second batch appends behind the first; the PLY file has the reference's header and 24 bytes per vertex
    saver.add_depthmap(t["inv_depth"].to(dev), t["image"].to(dev), t["K"].to(dev), t["pose"].to(dev), keep_masks=keeps if vote else ())
    assert len(saver) == 2 * ref.shape[0] and torch.equal(saver.vertices[ref.shape[0]:].cpu(), v)
    buf = io.BytesIO()
    saver.save(buf)
    raw = buf.getvalue()
    head, body = raw.split(b"end_header\n", 1)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % (2 * ref.shape[0]))
    assert len(body) == 2 * ref.shape[0] * 24


@pytest.mark.gpu
def test_cuda_pointcloud_dropout_and_growth():
    from monorec_b200 import pointcloud as PC
    from oracle import pointcloud_oracle as PO
    gen = torch.Generator().manual_seed(3)
    B, H, W = 3, 96, 160
    inv = torch.rand(B, 1, H, W, generator=gen) * 0.2 + 0.01
    img = torch.rand(B, 3, H, W, generator=gen) - 0.5
    K = torch.eye(4).repeat(B, 1, 1)
    K[:, 0, 0] = 150.0; K[:, 1, 1] = 149.0; K[:, 0, 2] = 80.0; K[:, 1, 2] = 48.0
    pose = torch.eye(4).repeat(B, 1, 1)
    pose[:, :3, 3] = torch.rand(B, 3, generator=gen) * 5
    rand = torch.rand(B, 1, H, W, generator=gen)
    ref = PO.add_depthmap(inv, img, K, pose, min_d=5, max_d=60, dropout=0.75, rand=rand)
    saver = PC.PLYSaver(H, W, min_d=5, max_d=60, batch_size=B, dropout=0.75)
    saver._buf = torch.empty(16, 6, device="cuda:0")               # force the growth path
    saver._count = torch.zeros(1, dtype=torch.int64, device="cuda:0")
    saver.add_depthmap(inv.cuda(), img.cuda(), K.cuda(), pose.cuda(), rand=rand.cuda())
    v = saver.vertices.cpu()
    assert v.shape == ref.shape and torch.allclose(v, ref, rtol=2e-6, atol=2e-5)
