"""The batched (baddbmm) critic ensemble is the same function as the per-member MLPEnsemble (franQ/Agent/models/mlp.py:64-108):
outputs and gradients agree to fp32 rounding, and the per-member state dicts round-trip.  CPU only."""
import pytest
import torch

from fastdeepqlearning_b200.Agent.components import models as m


@pytest.mark.parametrize("hidden", [(32, 16), (24,), (8, 8, 8)])
def test_batched_ensemble_equals_member_ensemble(hidden):
    torch.manual_seed(0)
    a, b = m.MLPEnsemble(20, 7, hidden, 3), m.BatchedMLPEnsemble(20, 7, hidden, 3)
    b.load_member_state_dicts([n.state_dict() for n in a.nets])
    x = torch.randn(4, 5, 20, requires_grad=True)
    ya, yb = a(x), b(x)
    assert ya.shape == yb.shape == (4, 5, 21)
    torch.testing.assert_close(ya, yb, rtol=1e-5, atol=1e-5)
    ga = torch.autograd.grad(ya.square().sum(), [x, a.nets[1].hidden[0].weight, a.nets[2].head.weight, a.nets[0].head.bias])
    gb = torch.autograd.grad(yb.square().sum(), [x, b.hidden_w[0], b.head_w, b.head_b])
    torch.testing.assert_close(ga[0], gb[0], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ga[1], gb[1][1].t(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ga[2], gb[2][2].t(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ga[3], gb[3][0, 0], rtol=1e-4, atol=1e-4)
    for sd, net in zip(b.member_state_dicts(), a.nets):
        for k, v in net.state_dict().items():
            assert torch.equal(sd[k], v), k
