"""Training-side kernels against stock torch: fused clamp+Adam over the flat bucket (train.py:62-65 semantics)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda:0')


def test_flat_adam_matches_torch_adam_with_clamp_hooks():
    from armnet_b200.parallel import FlatAdam
    torch.manual_seed(5)

    def make():
        torch.manual_seed(5)
        return nn.Sequential(nn.Linear(37, 19), nn.ReLU(), nn.Linear(19, 3)).to(dev())

    ref, ours = make(), make()
    opt = torch.optim.Adam(ref.parameters(), lr=3e-3)                     # train.py:62
    for p in ref.parameters():
        p.register_hook(lambda g: g.clamp(-1.0, 1.0))                     # train.py:64-65
    fa = FlatAdam(ours.parameters(), lr=3e-3, clamp=1.0)
    assert fa.numel() % 4 == 0
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        x = (torch.randn(64, 37, generator=g) * 30).to(dev())            # large inputs: the clamp is active
        y = torch.randn(64, 3, generator=g).to(dev())
        opt.zero_grad()
        ((ref(x) - y) ** 2).sum().backward()
        opt.step()
        fa.zero_grad()
        ((ours(x) - y) ** 2).sum().backward()
        fa.step()
        for a, b in zip(ref.parameters(), ours.parameters()):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (step, (a - b).abs().max().item())


def test_flat_adam_trains_the_drop_in_model():
    """FlatAdam re-points parameters to views of one buffer: the fused forward / backward must keep working on them."""
    import armnet_b200 as ab
    from armnet_b200.parallel import FlatAdam
    torch.manual_seed(0)
    F, V, B = 10, 500, 128
    model = ab.ARMNetModel(F, V, 16, 2, 1.7, 8, 1, 16, 0.0, False, 1, 8).to(dev()).train()
    fa = FlatAdam(model.parameters(), lr=1e-2, clamp=1.0)
    ids = torch.randint(0, V, (B, F), device=dev())
    target = (torch.rand(B, device=dev()) < 0.3).float()
    crit = nn.BCEWithLogitsLoss()
    losses = []
    for _ in range(30):
        fa.zero_grad()
        loss = crit(model({'id': ids, 'value': torch.ones(B, F, device=dev())}).reshape(-1), target)
        loss.backward()
        fa.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0]
