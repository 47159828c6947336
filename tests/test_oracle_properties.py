"""Property tests of the oracle itself, hypothesis-driven on tiny problems (CPU)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from monorec_b200.synthetic import make_inputs
from oracle import cost_volume_oracle as O


@settings(max_examples=6, deadline=None)
@given(seed=st.integers(0, 10_000), frames=st.integers(1, 3), planes=st.sampled_from([4, 8, 16]))
def test_ring_range_and_permutation(seed, frames, planes):
    data = make_inputs(2, frames, 24, 40, seed=seed)
    cv, sf = O.cost_volume_torch(data, steps=planes)
    for t in [cv] + sf:
        assert torch.isfinite(t).all() and t.abs().max() <= 1 + 1e-6
        assert (t[..., :2, :] == 0).all() and (t[..., -2:, :] == 0).all() and (t[..., :, :2] == 0).all() and (t[..., :, -2:] == 0).all()
    # batch independence: element 1 alone
    one = {k: ([t[1:2] for t in v] if isinstance(v, list) else v[1:2]) for k, v in data.items()}
    cv1, sf1 = O.cost_volume_torch(one, steps=planes)
    assert torch.allclose(cv1[0], cv[1], atol=1e-6) and all(torch.allclose(a[0], b[1], atol=1e-6) for a, b in zip(sf1, sf))
    if frames > 1:   # frame order: single-frame volumes permute, the fused volume is symmetric
        perm = list(reversed(range(frames)))
        pd = dict(data)
        for k in ("frames", "poses", "intrinsics"):
            pd[k] = [data[k][i] for i in perm]
        cvp, sfp = O.cost_volume_torch(pd, steps=planes)
        assert all(torch.allclose(sfp[j], sf[perm[j]], atol=1e-6) for j in range(frames))
        assert torch.allclose(cvp, cv, atol=1e-5)


@settings(max_examples=4, deadline=None)
@given(seed=st.integers(0, 10_000))
def test_closed_form_agrees_with_torch_restatement(seed):
    data = make_inputs(1, 2, 24, 40, seed=seed)
    cv, sf = O.cost_volume_torch(data, steps=8)
    cvc, sfc, valid, _ = O.cost_volume_closed_form(data, steps=8, dtype=np.float32)
    za, zb = (cv == 0).all(1), torch.from_numpy((cvc == 0).all(1))
    both = ~(za | zb)
    assert int((za != zb).sum()) <= 8
    assert ((cv - torch.from_numpy(cvc)).abs() * both.unsqueeze(1)).max() < 1e-3
