"""The oracle restatements vs the golden vectors minted from the reference itself (CPU, no GPU needed)."""
import numpy as np
import pytest
import torch

from oracle import cost_volume_oracle as O
from tests.helpers import compare_volumes, kitti_sample_dict, synth_small_dict


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])   # "d": 64 planes x 6 source frames (BASELINE config 5's counts)
def test_torch_restatement_matches_reference_small(tag):
    data, D, ref_cv, ref_sf = synth_small_dict(tag)
    cv, sf = O.cost_volume_torch(data, steps=D)
    # same primitives, same order: should be (nearly) bit-identical to the reference
    assert (cv - ref_cv).abs().max().item() <= 5e-5
    for a, r in zip(sf, ref_sf):
        assert (a - r).abs().max().item() <= 5e-5


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_closed_form_matches_reference_small(tag):
    data, D, ref_cv, ref_sf = synth_small_dict(tag)
    cv, sf, valid, _ = O.cost_volume_closed_form(data, steps=D, dtype=np.float32)
    stats = compare_volumes(torch.from_numpy(cv), [torch.from_numpy(s) for s in sf], ref_cv, ref_sf)
    print(tag, stats)


def test_torch_restatement_matches_reference_kitti():
    data, g = kitti_sample_dict()
    cv, sf = O.cost_volume_torch(data)
    sub = (slice(None), slice(None), slice(2, None, 4), slice(1, None, 8))
    # fp32 SSIM variance terms cancel catastrophically (SURVEY.md §7 hard part 2): a different BLAS summation
    # order in the projection already moves the volume by ~1e-4, so the gate is the north-star 1e-3, reported.
    d = np.abs(cv[sub].numpy() - g["cv_sub"]).max()
    print("kitti torch-restatement max|d| =", d)
    assert d <= 3e-4
    for i, v in enumerate(sf):
        assert np.abs(v[sub].numpy() - g["sf_sub"][i]).max() <= 3e-4
    assert (cv.argmax(1).numpy().astype(np.uint8) == g["argmax"]).mean() > 0.999
    np.testing.assert_allclose(cv.double().sum((2, 3)).numpy(), g["cv_plane_sum"], rtol=0, atol=5e-2)


def test_plane_depths_order():
    z = O.plane_depths(0.33, 0.0025, 32)
    assert abs(z[0].item() - 400.0) < 1e-2 and abs(z[-1].item() - 1 / 0.33) < 1e-5
    assert torch.all(z[1:] < z[:-1])


@pytest.mark.parametrize("gain_tag,gain", [("g1", 1.0), ("g07", 0.7)])
def test_convnet_oracle_matches_reference_model(gain_tag, gain):
    """oracle/convnet_oracle.py (+ the cost-volume oracle) vs the unmodified reference MonoRecModel forward."""
    from monorec_b200.model import MonoRecModel
    from monorec_b200.synthetic import make_inputs, seeded_state_dict
    from oracle import convnet_oracle as CO
    from tests.helpers import GOLDEN
    g = np.load(GOLDEN / "model_synth_small.npz")
    B, nF, D, H, W, seed, wseed = [int(v) for v in g["cfg"]]
    model = MonoRecModel()
    assert list(model.state_dict().keys()) == list(g["state_keys"])          # checkpoint contract, SURVEY.md §8b
    assert [",".join(str(int(v)) for v in t.shape) for t in model.state_dict().values()] == list(g["state_shapes"])
    sd = seeded_state_dict(model, seed=wseed, gain=gain)
    data = make_inputs(B, nF, H, W, seed=seed)
    cv, sf = O.cost_volume_torch(data, steps=D)
    out = CO.monorec_forward(sd, data, cv, sf)
    assert np.abs(out["cv_mask"].numpy() - g[f"{gain_tag}_cv_mask"]).max() < 2e-5
    for i, p in enumerate(out["predicted_inverse_depths"]):
        assert np.abs(p.numpy() - g[f"{gain_tag}_depth{i}"]).max() < 2e-5, i


@pytest.mark.parametrize("gain_tag,gain", [("g1", 1.0), ("g07", 0.7)])
def test_oracle_matches_reference_model_on_bundled_sample(gain_tag, gain):
    """The north-star parity sentence at the oracle level: full MonoRecModel on example/data's KITTI sample (256x512, 2 source
    frames, seeded weights) -- cost-volume oracle + conv-stack oracle vs the unmodified reference (model_kitti_sample.npz)."""
    from monorec_b200.model import MonoRecModel
    from monorec_b200.synthetic import seeded_state_dict
    from oracle import convnet_oracle as CO
    from tests.helpers import GOLDEN
    g = np.load(GOLDEN / "model_kitti_sample.npz")
    data, _ = kitti_sample_dict()
    cv, sf = O.cost_volume_torch(data)
    out = CO.monorec_forward(seeded_state_dict(MonoRecModel(), seed=int(g["wseed"][0]), gain=gain), data, cv, sf)
    assert np.abs(out["result"].numpy() - g[f"{gain_tag}_result"]).max() < 2e-5
    assert np.abs(out["cv_mask"].numpy() - g[f"{gain_tag}_cv_mask"].astype(np.float32)).max() < 6e-4   # stored as fp16
    for i, p in enumerate(out["predicted_inverse_depths"]):
        if i > 0:
            assert np.abs(p.numpy() - g[f"{gain_tag}_depth{i}"]).max() < 2e-5, i
