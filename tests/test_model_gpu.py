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
torch.backends.cudnn.allow_tf32 = False      # only matters for the "torch" comparison back-end
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(autouse=True)
def _tc_backend():
    from fsnet_b200.networks import ops
    ops.set_backend("tc")
    yield
    ops.set_backend("tc")


def to_cuda(data):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in data.items()}


@pytest.mark.parametrize("name", sorted(FULL_CASES))
def test_training_forward_backward_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    topo, B = FULL_CASES[name]["topo"], FULL_CASES[name]["B"]
    data = (O.synthetic_fisheye_batch if topo.fisheye else O.synthetic_batch)(B, topo.height, topo.width, 1234, topo.frame_ids)
    model = build_model(topo).cuda()
    model.head.tie_break_noise = O.tie_break_noise(B, topo.height, topo.width, topo.scales, 0)
    # piecewise, to see the maps (same orchestration as forward_train); on the tcgen05 path `feats` is a deferred
    # handle and forward_depth runs encoder + decoder as one autograd node
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
    floor = 1e-9 + 1e-7 * max(gn.values())       # conv biases in front of a BatchNorm: exactly zero here, rounding noise in the reference
    for k, p in model2.named_parameters():
        if k in gn and gn[k] > floor:
            e = abs(float(p.grad.double().norm()) - gn[k]) / gn[k]
            if e > (0.2 if topo.depth >= 50 else 0.1):      # gradients run through bf16 tensor-core operands
                bad.append((k, e))
    assert not bad, bad[:5]
    # eval-mode prediction after one train-mode forward (running statistics updated once, as in the fixture)
    model.eval()
    with torch.no_grad():
        pred = model(to_cuda(data), dict(is_training=False))
    assert rel(pred["depth"].cpu(), g["test_depth"]) < 2e-3
    if topo.fisheye:       # FishEyeDecoder.get_prediction: depth = z of the ray, plus the norm
        assert rel(pred["norm"].cpu(), g["test_norm"]) < 2e-3


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


def test_lazy_features_materialise_and_match_torch_backend():
    """`backbone(img)` is a drop-in list of five [B,C,h,w] tensors even on the tcgen05 path; the decoder still runs the real
    (differentiable) network when handed features somebody has already looked at (it used to fall back to eager ops
    on detached copies)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import torch_reference_backend
    topo = O.Topology(height=64, width=128)
    data = O.synthetic_batch(2, 64, 128, 1234)
    img = data[("image", 0)].cuda()
    model = build_model(topo).cuda()
    feats = model.depth_backbone(img)
    assert len(feats) == 5
    got = [f for f in feats]
    outs = model.head.forward_depth(feats, data["P2"].cuda())            # AFTER the features were materialised
    outs[("disp", 0)].sum().backward()
    assert model.depth_backbone.conv1.weight.grad is not None and float(model.depth_backbone.conv1.weight.grad.abs().sum()) > 0
    torch_reference_backend.enable()
    try:
        ref = build_model(topo).cuda().depth_backbone(img)
    finally:
        torch_reference_backend.disable()
    for a, b in zip(got, ref):
        assert a.shape == b.shape and rel(a.cpu(), b.detach().cpu()) < 1e-4
    with pytest.raises(RuntimeError):
        build_model(topo).depth_backbone(data[("image", 0)])              # CPU tensors: no eager fallback


def test_graphed_hook_matches_eager_hook():
    """CUDA-graph replay of the step == the eager step (same batch sequence, same Adam trajectory)."""
    from vision_base.utils.builder import build
    topo = O.Topology(height=64, width=128)
    losses = {}
    for mode in (False, True):
        torch.manual_seed(0)
        model = build_model(topo).cuda()
        # device-resident noise: a host->device copy cannot be captured into the step graph
        model.head.tie_break_noise = {s: n.cuda() for s, n in O.tie_break_noise(2, 64, 128, topo.scales, 0).items()}
        hook = build("vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=35.0, cuda_graph=mode)
        from vision_base.networks.optimizers.optimizers import build_optimizer
        opt = build_optimizer(model, name="adam", lr=1e-4) if mode else torch.optim.Adam(model.parameters(), lr=1e-4)   # FusedAdam vs stock
        seq = []
        for step in range(7):
            out = hook(O.synthetic_batch(2, 64, 128, 1234 + step), model, opt, None, None, step, 0)
            seq.append(float(out["loss"].detach()))
        losses[mode] = seq
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 2e-3 * abs(a), (losses[False], losses[True])


def test_fisheye_graphed_hook_matches_eager_hook():
    """FishEyeDecoder through BaseTrainingHook: the MEI ray table is (re)built on the device, so the step --
    table refresh included -- captures into one CUDA graph and replays the eager trajectory; `hm` images are
    produced because is_log_image is unset (=> True) in the shipped fisheye config."""
    from vision_base.utils.builder import build
    topo = O.Topology(height=64, width=64, fisheye=True, n_bins=64, max_depth=150.0)
    losses = {}
    for mode in (False, True):
        torch.manual_seed(0)
        model = build_model(topo, is_log_image=True).cuda()
        model.head.tie_break_noise = {s: n.cuda() for s, n in O.tie_break_noise(2, 64, 64, topo.scales, 0).items()}
        hook = build("vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=1.0, cuda_graph=mode)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        seq = []
        for step in range(7):
            # the calibration changes from batch to batch (KITTI360FisheyeDataset draws image_02 / image_03 per sample): replays must
            # see it (the list-of-dicts `calib_meta` was frozen at its capture value)
            out = hook(O.synthetic_fisheye_batch(2, 64, 64, 1234 + step, two_calibrations=(step % 2 == 1)), model, opt, None, None, step, 0)
            seq.append(float(out["loss"].detach()))
        assert out["hm"]["predicted_image_1"].shape == (1, 3, 64, 64) and out["hm"]["loss_mask_0"]["data"].dtype == torch.bool
        losses[mode] = seq
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 2e-3 * abs(a), (losses[False], losses[True])
