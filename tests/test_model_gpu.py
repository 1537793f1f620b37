"""GPU parity of the whole training forward / backward through the reference-compatible plugin path,
against the golden fixtures (outputs of the reference itself).  Tolerance: north_star's 1e-3 relative
on disparity maps and loss scalars."""
import numpy as np
import pytest
import torch

from oracle import fsnet_oracle as O
from helpers import build_model
from test_oracle_golden import FULL_CASES, rel, load

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False      # interim library convolutions must meet the same 1e-3 bar
torch.backends.cuda.matmul.allow_tf32 = False


def to_cuda(data):
    return {k: v.cuda() for k, v in data.items()}


@pytest.mark.parametrize("name", sorted(FULL_CASES))
def test_training_forward_backward_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    topo, B = FULL_CASES[name]["topo"], FULL_CASES[name]["B"]
    data = O.synthetic_batch(B, topo.height, topo.width, 1234, topo.frame_ids)
    model = build_model(topo).cuda()
    model.head.tie_break_noise = O.tie_break_noise(B, topo.height, topo.width, topo.scales, 0)
    # piecewise, to see the maps (same orchestration as forward_train)
    feats = model.depth_backbone(data[("image", 0)].cuda())
    outs = model.head.forward_depth(feats) if topo.posenet else model.head.forward_depth(feats, data["P2"].cuda())
    for s in topo.scales:
        assert rel(outs[("disp", s)].detach().cpu(), g[f"disp/{s}"]) < 1e-3, s
        assert rel(outs[("depth", s, s)].detach().cpu(), g[f"depth/{s}"]) < 1e-3, s
    # the public entry
    model2 = build_model(topo).cuda()
    model2.head.tie_break_noise = model.head.tie_break_noise
    ret = model2(to_cuda(data), dict(is_training=True, epoch_num=0, global_step=0))
    assert (ret["loss"].dtype == torch.float64) == bool(g["loss_is_fp64"])
    assert abs(float(ret["loss"].detach()) - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    for k, v in ret["loss_dict"].items():
        ref = float(g["loss_dict/" + k])
        assert abs(float(v) - ref) <= 1e-3 * abs(ref) + 1e-12, (k, float(v), ref)
    ret["loss"].mean().backward()
    gn = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    bad = []
    for k, p in model2.named_parameters():
        if k in gn and gn[k] > 1e-9:
            e = abs(float(p.grad.double().norm()) - gn[k]) / gn[k]
            if e > 5e-2:
                bad.append((k, e))
    assert not bad, bad[:5]
    # eval-mode prediction after one train-mode forward (running statistics updated once, as in the fixture)
    model.eval()
    with torch.no_grad():
        pred = model(to_cuda(data), dict(is_training=False))
    assert rel(pred["depth"].cpu(), g["test_depth"]) < 2e-3


def test_training_hook_steps_and_loss_decreases():
    """BaseTrainingHook drives the model exactly like the reference loop; a few Adam steps on one batch
    must reduce the loss (end-to-end sanity of every backward kernel)."""
    from vision_base.utils.builder import build
    topo = O.Topology(height=64, width=128)
    model = build_model(topo).cuda()
    hook = build("vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=35.0)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for step in range(12):
        data = O.synthetic_batch(2, 64, 128, 1234)
        out = hook(data, model, opt, None, None, step, 0)
        losses.append(float(out["loss"].detach()))
    assert np.isfinite(losses).all()
    assert min(losses[-4:]) < losses[0], losses
