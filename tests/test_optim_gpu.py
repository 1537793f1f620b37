"""FusedAdam (fsnet_grad_sumsq + fsnet_adam_step) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("clip,wd", [(35.0, 0.0), (0.05, 0.0), (None, 1e-2)])
def test_fused_adam_matches_torch(clip, wd):
    from fsnet_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(64, 3, 7, 7), (64,), (5,), (128, 64, 3, 3), (1,), (4097,), (16, 16, 3, 3), (12289,)]
    ref = [torch.randn(s, device="cuda", generator=g).requires_grad_(True) for s in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-3, weight_decay=wd)
    o_mine = FusedAdam(mine, lr=1e-3, weight_decay=wd)
    sched = torch.optim.lr_scheduler.StepLR(o_mine, step_size=3, gamma=0.5)
    sched_ref = torch.optim.lr_scheduler.StepLR(o_ref, step_size=3, gamma=0.5)
    for step in range(8):
        for a, b in zip(ref, mine):
            a.grad = torch.randn(a.shape, device="cuda", generator=g) * (0.1 + step)
            b.grad = a.grad.clone()
        if clip is not None:
            norm_ref = torch.nn.utils.clip_grad_norm_(ref, clip)
        o_ref.step()
        o_mine.step(max_norm=clip)
        if clip is not None:
            assert abs(float(o_mine.total_norm()) - float(norm_ref)) <= 1e-5 * float(norm_ref)
        sched.step(); sched_ref.step()
    for a, b in zip(ref, mine):
        assert float((a - b).abs().max()) <= 2e-6 * (1 + float(a.abs().max())), float((a - b).abs().max())
    # same state_dict layout as torch.optim.Adam; a reloaded optimiser continues the trajectory
    sd = o_mine.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 8
    o2 = FusedAdam(mine, lr=1e-3, weight_decay=wd)
    o2.load_state_dict(sd)
    o_ref.param_groups[0]["lr"] = o2.param_groups[0]["lr"]
    for a, b in zip(ref, mine):
        a.grad = torch.randn(a.shape, device="cuda", generator=g)
        b.grad = a.grad.clone()
    o_ref.step(); o2.step()
    for a, b in zip(ref, mine):
        assert float((a - b).abs().max()) <= 2e-6 * (1 + float(a.abs().max()))
