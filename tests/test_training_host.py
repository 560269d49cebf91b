"""CPU tests of the host-side training logic (no GPU): packed <-> PyTorch parameter layouts, RLOO advantages, the
trainer refusing to run without CUDA."""
import pytest
import torch

from oracle import sd3_oracle as O
from tpdm_b200.rloo import rloo_advantages
from tpdm_b200.tpm_training import NAMES, _from_packed, _to_packed


def test_packed_layout_round_trip_and_meaning():
    torch.manual_seed(0)
    tp = O.OracleTimePredictor(128, 256)
    sd = tp.state_dict()
    assert set(NAMES) == set(sd)
    for name in NAMES:
        flat = _to_packed(name, sd[name])
        assert flat.numel() == sd[name].numel()
        assert torch.equal(_from_packed(name, flat, sd[name]), sd[name])
    w1 = sd["conv1.weight"]                                  # [C1, 2D, 3, 3] -> [C1][9][2D], tap = ky*3 + kx
    p1 = _to_packed("conv1.weight", w1).reshape(128, 9, 256)
    assert torch.equal(p1[5, 2 * 3 + 1, 17], w1[5, 17, 2, 1])
    w2 = sd["conv2.weight"]                                  # [oc, c, 3, 3] -> [9][c][oc]
    p2 = _to_packed("conv2.weight", w2).reshape(9, 128, 128)
    assert torch.equal(p2[1 * 3 + 2, 40, 7], w2[7, 40, 1, 2])


def test_rloo_advantages_match_oracle_restatement():
    r = torch.tensor([0.3, -1.0, 2.0, 0.5, 1.5, -0.2, 0.0, 0.7])
    for k in (2, 4):
        assert torch.allclose(rloo_advantages(r, k), O.rloo_advantage(r, k))
    assert torch.allclose(rloo_advantages(r, 4).reshape(4, -1).sum(0), torch.zeros(2), atol=1e-6)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_trainer_requires_cuda():
    from tpdm_b200.modeling_sd3_pnt import TimePredictor
    from tpdm_b200.tpm_training import TimePredictorTrainer

    with pytest.raises(RuntimeError):
        TimePredictorTrainer(TimePredictor(128, 256), grid=16, max_samples=2)
