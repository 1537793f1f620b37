"""The loss-side CUDA kernels themselves, in the CPU suite: fsnet_b200/csrc/{warp_ssim,smooth_head,optim,distill,augment}.cu are
compiled for the host on top of a small SIMT runtime (tests/host_emulation/simt.h: one fiber per CUDA thread, warp shuffles /
__syncthreads as barriers between fibers, blocks run in sequence) and driven through the SAME C ABI and the SAME Python
autograd Functions as on the GPU -- against the golden vectors produced by the reference.

What this proves: index arithmetic, warp-level data exchange, reductions, branch logic and the host-side launch planning of those
kernels.  What it can not: anything about speed, memory-ordering bugs that need real parallelism, and the tcgen05 / TMA kernels
(convolutions, BatchNorm planes) -- those stay with the -m gpu suite.  FSNET_EMULATE_ALL=1 adds the remaining model families (~2 more minutes)."""
import os

import numpy as np
import pytest
import torch

from oracle import fsnet_oracle as O
from test_loss_gpu import depth_grad_ok, run_gpu_loss
from test_oracle_golden import LOSS_CASES, build_loss_case, load, rel

CASES = sorted(LOSS_CASES)


@pytest.fixture
def emulated(monkeypatch):
    from host_emulation import fixture
    return fixture.install(monkeypatch)


@pytest.mark.parametrize("name", CASES)
def test_fused_loss_kernels_match_golden_under_emulation(emulated, golden_dir, name):
    """Same assertions as tests/test_loss_gpu.py::test_fused_loss_matches_golden (pinhole, motion mask, MEI fisheye incl. the
    fp64 ray-table build; forward, fused forward+backward, pose gradients)."""
    g = load(golden_dir, name)
    case = LOSS_CASES[name]
    topo = case["topo"]
    data, outputs, noise = build_loss_case(**case)
    total, stats, depths, disps, T = run_gpu_loss(topo, data, outputs, noise, dev="cpu")
    S = len(topo.scales)
    assert abs(float(total.detach()) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    assert (total.dtype == torch.float64) == ("patched_mask" in data and data["patched_mask"].dtype == torch.float64)
    for i, s in enumerate(topo.scales):
        assert abs(float(stats[i]) - float(g[f"loss_dict/loss/{s}"])) <= 1e-4 * abs(float(g[f"loss_dict/loss/{s}"])), s
        assert abs(float(stats[S + i]) - float(g[f"loss_dict/smooth_loss/{s}"])) <= 1e-4 * abs(float(g[f"loss_dict/smooth_loss/{s}"])), s
        assert rel(disps[i].grad, g[f"grad_disp/{s}"]) < 1e-3, s
        ok, e = depth_grad_ok(depths[i].grad, g[f"grad_depth/{s}"])
        assert ok, (s, e)
    for fi, f in enumerate(topo.frame_ids[1:]):
        assert rel(T[fi].grad, g[f"grad_T/{f}"]) < 0.3, f


@pytest.mark.parametrize("name", CASES)
def test_frame_pair_loss_kernel_matches_golden_under_emulation(emulated, golden_dir, name):
    """The fused forward+backward launch of a training step WITHOUT pose gradients (dataset poses: every shipped config) runs
    loss_pair_kernel (two warps per strip, one per source frame; halo of one pixel, partial gradients added in memory)."""
    g = load(golden_dir, name)
    case = LOSS_CASES[name]
    topo = case["topo"]
    data, outputs, noise = build_loss_case(**case)
    total, stats, depths, disps, T = run_gpu_loss(topo, data, outputs, noise, need_pose=False, dev="cpu")
    S = len(topo.scales)
    assert abs(float(total.detach()) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    for i, s in enumerate(topo.scales):
        assert abs(float(stats[i]) - float(g[f"loss_dict/loss/{s}"])) <= 1e-4 * abs(float(g[f"loss_dict/loss/{s}"])), s
        ok, e = depth_grad_ok(depths[i].grad, g[f"grad_depth/{s}"])
        assert ok, (s, e)


@pytest.mark.parametrize("H,W,B,scales,fisheye", [(40, 72, 2, (0, 1, 2), False), (24, 61, 1, (0,), False), (24, 88, 1, (0, 3), False), (56, 104, 1, (0, 1), True),
                                                  (9, 31, 1, (0,), False), (16, 30, 2, (0, 1), False)])
def test_frame_pair_loss_kernel_on_ragged_shapes_under_emulation(emulated, H, W, B, scales, fisheye):
    """Widths around the 30-column strip (30, 31, 61), heights that leave short last chunks, single samples."""
    topo = O.Topology(height=H, width=W, scales=scales, fisheye=fisheye, max_depth=150.0 if fisheye else 100.0)
    data, outputs, noise = build_loss_case(topo, B, 31)
    for v in outputs.values():
        v.requires_grad_(True)
    cam_T = {f: data[("relative_pose", f)] for f in topo.frame_ids[1:]}
    ref = O.loss_chain(outputs, data, cam_T, topo, noise, keep=True)
    ref["loss"].backward()
    total, stats, depths, disps, T = run_gpu_loss(topo, data, outputs, noise, need_pose=False, dev="cpu")
    assert abs(float(total.detach()) - float(ref["loss"].detach())) <= 1e-5 * abs(float(ref["loss"].detach()))
    for i, s in enumerate(scales):
        ok, e = depth_grad_ok(depths[i].grad, outputs[("depth", s, s)].grad)
        assert ok, (s, e)


def test_log_image_head_keeps_the_loss_and_gradients_under_emulation(emulated, golden_dir):
    """the patched-mask normaliser of a non-fused scale 0 (log-image head) must not be counted twice."""
    from test_loss_gpu import check_log_image_case
    check_log_image_case(golden_dir, "loss_a", "cpu")


def test_pose_matrix_kernels_under_emulation(emulated):
    """fsnet_pose_matrix / _bwd (axis-angle + translation -> 4x4, forward-mode duals contracted with the incoming gradient) against
    autograd through the oracle's transformation_from_parameters, both orientations, incl. a near-zero rotation."""
    from fsnet_b200 import functional as Fn
    g = torch.Generator().manual_seed(4)
    for invert in (False, True):
        aa = (torch.randn(5, 1, 3, generator=g) * 0.05)
        aa[0] = 1e-6 * torch.randn(1, 3, generator=g)
        tr = torch.randn(5, 1, 3, generator=g)
        gT = torch.randn(5, 4, 4, generator=g)
        a1, t1 = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
        ref = O.transformation_from_parameters(a1, t1, invert)
        (ref * gT).sum().backward()
        a2, t2 = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
        got = Fn.pose_matrix(a2, t2, invert)
        (got * gT).sum().backward()
        assert float((got - ref).abs().max()) < 1e-6
        assert rel(a2.grad, a1.grad) < 1e-5 and rel(t2.grad, t1.grad) < 1e-6


def test_depth_head_kernels_under_emulation(emulated):
    """fsnet_depth_head_fwd / _bwd (channels-last lane-group softmax and the NCHW kernel) against the oracle's gather_depth."""
    from fsnet_b200 import functional as Fn
    g = torch.Generator().manual_seed(2)
    for n, channels_last in ((16, True), (64, True), (16, False)):
        topo = O.Topology(n_bins=n, base_fx=40.0)
        logits = (torch.randn(2, n, 6, 10, generator=g) * 4)
        if channels_last:
            logits = logits.contiguous(memory_format=torch.channels_last)
        logits.requires_grad_(True)
        bins = O.depth_bins(topo)
        scale = torch.tensor([0.8, 1.3])
        depth, disp = Fn.depth_head(logits, bins, scale, False, topo.min_depth, topo.max_depth)
        (depth.sum() + 3 * disp.sum()).backward()
        ref_logits = logits.detach().clone().requires_grad_(True)
        d_ref, disp_ref = O.gather_depth(ref_logits, bins, topo, scale.reshape(-1, 1, 1, 1))
        (d_ref.sum() + 3 * disp_ref.sum()).backward()
        assert rel(depth, d_ref) < 1e-5 and rel(disp, disp_ref) < 1e-5
        assert rel(logits.grad, ref_logits.grad) < 1e-4


def test_distill_loss_kernel_under_emulation(emulated):
    """fsnet_distill_loss, the whole kernel (grid-stride loop, warp + block reduction, atomic accumulation) through
    functional.distill_loss, against torch autograd on the reference's expression."""
    from fsnet_b200 import functional as Fn
    g = torch.Generator().manual_seed(3)
    for n, with_u in ((5, True), (3000, True), (70001, False)):
        p = (torch.rand(n, generator=g) * 40 + 1).requires_grad_(True)
        t = torch.rand(n, generator=g) * 40 + 1
        t[: n // 7] = p.detach()[: n // 7]
        l = (torch.randn(n, generator=g) * 2).requires_grad_(True) if with_u else None
        out = Fn.distill_loss(p.view(1, 1, 1, n), t.view(1, 1, 1, n), None if l is None else l.view(1, 1, 1, n))
        (out * 0.3).backward()
        p2 = p.detach().clone().requires_grad_(True)
        l2 = None if l is None else l.detach().clone().requires_grad_(True)
        err = (t - p2).abs()
        ref = (err / torch.sigmoid(l2) + torch.log(torch.sigmoid(l2) + 1e-5)).mean() if with_u else err.mean()
        (ref * 0.3).backward()
        assert abs(float(out.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach())) + 1e-7
        assert rel(p.grad, p2.grad) < 1e-4
        if with_u:
            assert rel(l.grad, l2.grad) < 1e-4


def test_device_augmentation_stage_under_emulation(emulated, golden_dir):
    """DeviceAugmentStage -> fsnet_augment_frames through the C ABI: the reference pipeline's golden vectors."""
    from aug_cases import raw_sample
    from test_device_augment_cpu import device_cfg
    from fsnet_b200.data.device_augment import DeviceAugmentStage, device_augment_collate
    from vision_base.utils.builder import build
    g = np.load(os.path.join(golden_dir, "aug_train.npz"))
    np.random.seed(7)
    aug = build(**device_cfg())
    batch = DeviceAugmentStage(aug)(device_augment_collate([aug(raw_sample(100 + i)) for i in range(3)]))
    for i in range(3):
        np.testing.assert_allclose(batch[("image", 0)][i].numpy(), g[f"{i}/full/image_0"], rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(batch[("original_image", 1)][i].numpy(), g[f"{i}/full/original_image_1"], rtol=1e-5, atol=2e-5)
    assert batch["patched_mask"].dtype == torch.float64 and batch["patched_mask"].shape == (3, 48, 160)


@pytest.mark.parametrize("clip,wd", [(35.0, 0.0), (0.05, 0.0), (None, 1e-2)])
def test_fused_adam_kernels_under_emulation(emulated, clip, wd):
    """fsnet_grad_sumsq + fsnet_adam_step (device tensor table, on-the-fly clip, device-resident step counter) against
    clip_grad_norm_ + torch.optim.Adam -- the trajectory test of tests/test_optim_gpu.py."""
    from fsnet_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(8, 3, 7, 7), (64,), (5,), (16, 8, 3, 3), (1,), (4097,), (9000,)]
    ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-3, weight_decay=wd)
    o_mine = FusedAdam(mine, lr=1e-3, weight_decay=wd)
    sched, sched_ref = (torch.optim.lr_scheduler.StepLR(o, step_size=3, gamma=0.5) for o in (o_mine, o_ref))
    for step in range(6):
        for a, b in zip(ref, mine):
            a.grad = torch.randn(a.shape, generator=g) * (0.1 + step)
            b.grad = a.grad.clone()
        if clip is not None:
            norm_ref = torch.nn.utils.clip_grad_norm_(ref, clip)
        o_ref.step()
        o_mine.step(max_norm=clip)
        if clip is not None:
            assert abs(float(o_mine.total_norm()) - float(norm_ref)) <= 1e-5 * float(norm_ref)
        sched.step()
        sched_ref.step()
    for a, b in zip(ref, mine):
        assert float((a - b).abs().max()) <= 2e-6 * (1 + float(a.abs().max()))
    assert float(o_mine.state_dict()["state"][0]["step"]) == 6


FULL = ["micro_distill", "micro_normeval_frozen", "tiny4", "tiny_pose", "tiny_fe"] + (
    ["tiny_sigmoid", "tiny_r50", "tiny_distill", "tiny_normeval", "tiny_frozen", "tiny_normeval_frozen"]
    if os.environ.get("FSNET_EMULATE_ALL") == "1" else [])


@pytest.mark.parametrize("name", FULL)
def test_whole_training_step_through_the_executor_under_emulation(emulated, golden_dir, name):
    """A whole training step on the CPU THROUGH THE tcgen05 EXECUTOR PATH (fsnet_b200/engine.py: planes, tape, BatchNorm /
    activation / pooling / re-layout kernels of act_tc.cu emulated; the two tensor-core entry points replaced by the ABI-level
    stand-ins of tests/host_emulation/conv_ref.cpp) plus the emulated loss side, against the golden produced by the reference.
    Default cases: the second-stage model (DistillWPoseMeta: frozen eval-mode teacher, student with the uncertainty heads as a
    second convolution on the decoder activations, distillation loss), ResNet(norm_eval, frozen_stages), the cfg2a topology, the
    PoseNet variant and the fisheye head -- loss_dict, total loss, gradient norms, sample gradients, exactly the reference's set
    of trained parameters, untouched teacher."""
    from helpers import build_model
    from fsnet_b200.networks import ops
    from test_oracle_golden import ALL_FULL_CASES
    case = ALL_FULL_CASES[name]
    topo, B = case["topo"], case["B"]
    g = load(golden_dir, name)
    data = (O.synthetic_fisheye_batch if topo.fisheye else O.synthetic_batch)(B, topo.height, topo.width, 1234, topo.frame_ids)
    assert ops.tc_available()
    backend = "tc"
    ops.set_backend("tc")
    try:
        model = build_model(topo)
        model.head.tie_break_noise = O.tie_break_noise(B, topo.height, topo.width, topo.scales, 0)
        teacher_before = {k: v.clone() for k, v in model.teacher_net.state_dict().items()} if topo.distill else {}
        ret = model(dict(data), dict(is_training=True, epoch_num=0, global_step=0))
        ret["loss"].mean().backward()
    finally:
        ops.set_backend(backend)
    assert (ret["loss"].dtype == torch.float64) == bool(g["loss_is_fp64"])
    assert abs(float(ret["loss"].detach()) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    for k, v in ret["loss_dict"].items():
        ref = float(g["loss_dict/" + k])
        assert abs(float(v) - ref) <= 1e-3 * abs(ref) + 1e-12, (k, float(v), ref)
    gn = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    floor = 1e-9 + 1e-7 * max(gn.values())          # conv biases in front of a BatchNorm: exactly zero here, rounding noise there
    params = dict(model.named_parameters())
    assert {k for k, p in params.items() if p.grad is not None} == set(gn)      # exactly the parameters the reference trains
    for k, p in params.items():
        if k.startswith("teacher_net."):
            assert p.grad is None and k not in gn
        elif k in gn:
            assert abs(float(p.grad.double().norm()) - gn[k]) <= 0.05 * gn[k] + floor, (k, float(p.grad.double().norm()), gn[k])
    for key in g.files:
        if key.startswith("grad/"):
            # gradients run through bf16 operands (hi planes only); at 32x64 the deepest maps are 1x2 pixels and the BatchNorm
            # batches 4 values, which amplifies that rounding
            assert rel(params[key[5:]].grad, g[key]) < (0.15 if name.startswith("micro") else 0.05), key
    if topo.distill:
        assert {f"distilation/{s}" for s in topo.scales} <= set(ret["loss_dict"])
        assert all(torch.equal(teacher_before[k], v) for k, v in model.teacher_net.state_dict().items())
    # the executor hands its gradients to autograd without keeping a reference: p.grad ARE the views of the per-network flat
    # buffers (plus the shared all-zero buffer of BatchNorm-cancelled conv biases), not one private clone per parameter --
    # which is what makes the data-parallel exchange a single in-place all-reduce (hooks/training.py::sync_gradients)
    # (the PoseNet runs twice per step with shared weights: autograd sums its two gradients into fresh tensors first)
    storages = {p.grad.untyped_storage().data_ptr() for k, p in params.items() if p.grad is not None and "pose" not in k}
    assert len(storages) <= 4, len(storages)


@pytest.mark.parametrize("H,W,B,scales,fisheye", [(40, 72, 2, (0, 1, 2), False), (24, 40, 1, (0, 3), False), (56, 104, 1, (0, 1), True)])
def test_fused_loss_on_ragged_shapes_under_emulation(emulated, H, W, B, scales, fisheye):
    """Image sizes that are no multiple of the warp width or of the row tiles, a single sample, scale subsets: the marching-warp
    kernels against the oracle's loss chain (value 1e-5, disparity gradients 1e-4, depth gradients up to arg-min ties)."""
    topo = O.Topology(height=H, width=W, scales=scales, fisheye=fisheye, max_depth=150.0 if fisheye else 100.0)
    data, outputs, noise = build_loss_case(topo, B, 31)
    for v in outputs.values():
        v.requires_grad_(True)
    cam_T = {f: data[("relative_pose", f)].clone().requires_grad_(True) for f in topo.frame_ids[1:]}
    ref = O.loss_chain(outputs, data, cam_T, topo, noise, keep=True)
    ref["loss"].backward()
    total, stats, depths, disps, T = run_gpu_loss(topo, data, outputs, noise, dev="cpu")
    assert abs(float(total.detach()) - float(ref["loss"].detach())) <= 1e-5 * abs(float(ref["loss"].detach()))
    for i, s in enumerate(scales):
        assert rel(disps[i].grad, outputs[("disp", s)].grad) < 1e-4, s
        ok, e = depth_grad_ok(depths[i].grad, outputs[("depth", s, s)].grad.numpy())
        assert ok, (s, e)


def test_training_hook_trains_the_distillation_model_under_emulation(emulated, monkeypatch):
    """BaseTrainingHook (eager) + FusedAdam over model.parameters() of DistillWPoseMeta: the loss goes down, the frozen teacher is
    never touched, optimiser state exists for the trainable parameters only."""
    from helpers import build_model
    from fsnet_b200.networks import ops
    from fsnet_b200.optim import build_optimizer
    from vision_base.utils.builder import build
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    backend = "tc"
    ops.set_backend("tc")
    try:
        topo = O.Topology(height=32, width=64, distill=True)
        model = build_model(topo)
        opt = build_optimizer(model, name="adam", lr=1e-3, weight_decay=0)
        hook = build(name="vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=35.0, cuda_graph=False)
        data = O.synthetic_batch(2, 32, 64, 1234, topo.frame_ids)
        teacher = {k: v.clone() for k, v in model.teacher_net.state_dict().items()}
        losses = [float(hook(dict(data), model, opt, None, None, i, 0)["loss"].detach()) for i in range(3)]
    finally:
        ops.set_backend(backend)
    assert losses[-1] < losses[0]
    assert all(torch.equal(teacher[k], v) for k, v in model.teacher_net.state_dict().items())
    assert len(opt.state_dict()["state"]) == sum(p.requires_grad for p in model.parameters())


PLAN_CASES = {
    "cfg2a kitti 192x640 R18 B12": (O.Topology(height=192, width=640), 12),
    "cfg2b +PoseNet B12": (O.Topology(height=192, width=640, posenet=True, overlapped_mask=False), 12),
    "cfg3 192x768 R50 B8": (O.Topology(height=192, width=768, depth=50), 8),
    "cfg4 320x640 R18 B8": (O.Topology(height=320, width=640), 8),
    "nusc shipped 288x512 R34 n64 B8": (O.Topology(height=288, width=512, depth=34, n_bins=64, base_fx=369.0, overlapped_mask=False), 8),
    "cfg5 fisheye 512x512 n64 B4": (O.Topology(height=512, width=512, fisheye=True, n_bins=64, max_depth=150.0), 4),
    "fisheye shipped 384x384 B16": (O.Topology(height=384, width=384, fisheye=True, n_bins=64, max_depth=150.0), 16),
    "distillation 192x640 B12": (O.Topology(height=192, width=640, distill=True), 12),
    "R101 192x640 B4": (O.Topology(height=192, width=640, depth=101), 4),
}


@pytest.mark.parametrize("name", sorted(PLAN_CASES))
def test_conv_planner_accepts_every_shipped_configuration(emulated, monkeypatch, name):
    """Every tensor-core launch of a full training step (forward, data gradient, weight gradient; all layers) of the
    configurations BASELINE.json / the reference ship, at their real image and batch sizes, goes through conv_tc.cu's own
    HOST code -- argument checks, tile / pipeline-depth / K-split planning, tensor-map encoding against a validating
    cuTensorMapEncodeTiled (tests/host_emulation/cuda.h) -- with the kernels themselves skipped (values are garbage).
    Catches 'unsupported shape' / 'does not fit shared memory' / invalid tensor-map failures without a GPU."""
    import ctypes
    from helpers import build_model
    from fsnet_b200.networks import ops
    monkeypatch.setenv("FSNET_EMULATE_PLAN_ONLY", "1")
    topo, B = PLAN_CASES[name]
    data = (O.synthetic_fisheye_batch if topo.fisheye else O.synthetic_batch)(B, topo.height, topo.width, 1, topo.frame_ids)
    counter = ctypes.c_longlong.in_dll(emulated, "fsnet_emulated_plans")
    before = counter.value
    backend = "tc"
    ops.set_backend("tc")
    try:
        model = build_model(topo)
        model(dict(data), dict(is_training=True, epoch_num=0, global_step=0))["loss"].mean().backward()
    finally:
        ops.set_backend(backend)
    assert counter.value - before >= 100           # 34 convolutions x (forward, data gradient, weight gradient) for ResNet-18


def test_conv_planner_rejects_what_the_kernels_can_not_do(emulated):
    """The stand-in chain is not a rubber stamp: the planner's own checks fire through it."""
    from fsnet_b200 import _lib, tc
    x = tc.Planes(1, 8, 8, 24, ring=1, device="cpu", zero=True)            # 24 channels: not a multiple of 16
    w = torch.zeros(2, 16, 3, 3, 24, dtype=torch.bfloat16)
    out = tc.Fp32(1, 8, 8, 16, device="cpu")
    with pytest.raises(_lib.FsnetError, match="multiples of 16"):
        _lib.call("fsnet_conv", x.view(), 0, w[0], w[1], 16, 3, 3, 1, 1, 3, None, 0, out.view(), 0, None)
    x = tc.Planes(1, 8, 8, 16, ring=0, device="cpu", zero=True)
    with pytest.raises(_lib.FsnetError, match="ring"):                      # replicate padding without a materialised ring
        _lib.call("fsnet_conv", x.view(), 1, w[0, :, :, :, :16].contiguous(), w[1, :, :, :, :16].contiguous(), 16, 3, 3, 1, 1, 3, None, 0,
                  out.view(), 0, None)


@pytest.mark.parametrize("variant", ["plain"] + (["norm_eval_frozen"] if os.environ.get("FSNET_EMULATE_ALL") == "1" else []))
def test_training_trajectory_matches_the_reference_step_for_step(emulated, monkeypatch, variant):
    """Three consecutive steps of BaseTrainingHook (zero_grad, forward, backward, clip 35, FusedAdam) through the executor against
    the oracle's restatement of the reference's training step (torch Adam on the CPU): per-step losses to north_star's 1e-3,
    parameters and BatchNorm running statistics after the last step.  Exercises what a single step can not: the batched
    operand-plane refresh from updated weights, running-statistics updates feeding nothing but later eval, the optimiser state."""
    from helpers import build_model
    from fsnet_b200.networks import ops
    from fsnet_b200.optim import build_optimizer
    from vision_base.utils.builder import build
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    frozen = variant == "norm_eval_frozen"
    topo, B = O.Topology(height=32, width=64, norm_eval=frozen, frozen_stages=1 if frozen else -1), 2
    backend = "tc"
    ops.set_backend("tc")
    try:
        model = build_model(topo)
        before = {k: v.clone() for k, v in model.state_dict().items()}
        opt = build_optimizer(model, name="adam", lr=1e-4, weight_decay=0)
        hook = build(name="vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook", clip_gradients=35.0, cuda_graph=False)
        trainer = O.OracleTrainer(topo, lr=1e-4, clip=35.0)
        for i in range(3):
            data = O.synthetic_batch(B, topo.height, topo.width, 100 + i, topo.frame_ids)
            noise = O.tie_break_noise(B, topo.height, topo.width, topo.scales, i)
            model.head.tie_break_noise = noise
            mine = float(hook(dict(data), model, opt, None, None, i, 0)["loss"].detach())
            ref = float(trainer.step(data, noise)["loss"].detach())
            assert abs(mine - ref) <= 1e-3 * abs(ref), (i, mine, ref)
    finally:
        ops.set_backend(backend)
    sd = model.state_dict()
    if frozen:      # frozen stem + layer1 and every encoder running statistic are exactly what they were
        assert all(torch.equal(before[k], v) for k, v in sd.items()
                   if k.startswith(("depth_backbone.conv1", "depth_backbone.bn1", "depth_backbone.layer1")) or
                   (k.startswith("depth_backbone") and "running" in k))
    for k, v in sd.items():
        if frozen and k.endswith("num_batches_tracked") and k.startswith("depth_backbone"):
            assert int(v) == 0, k
            continue
        if v.is_floating_point() and not k.endswith("depth_bins"):
            assert rel(v, trainer.sd[k]) < 5e-2, k        # 1x2-pixel maps at this size: statistics over 4 values amplify the bf16x3 rounding
        elif k.endswith("num_batches_tracked"):
            assert int(v) == 3, k          # nn.BatchNorm2d's counter (the functional oracle does not keep one)


def test_files_to_training_step_with_device_augmentation_under_emulation(emulated, monkeypatch, tmp_path):
    """The whole input side in one chain: miniature KITTI tree -> reader with DeviceAugmentation (uint8 frames + drawn parameters)
    -> padded collate -> DevicePrefetcher with the device stage (fsnet_augment_frames) -> BaseTrainingHook step through the
    executor.  The batch the model sees has the reference pipeline's schema; two steps run and produce finite, different losses."""
    from kitti_fixture import build_tree
    from fsnet_b200.data.device_augment import device_augment_collate, find_device_stage
    from fsnet_b200.data.loading import DevicePrefetcher, build_dataloader
    from fsnet_b200.networks import ops
    from fsnet_b200.optim import build_optimizer
    from vision_base.utils.builder import build
    from vision_base.utils.utils import cfg_from_file
    raw, split = build_tree(str(tmp_path))
    for k, v in dict(FSNET_KITTI_PATH=raw, FSNET_KITTI_SPLIT=split, FSNET_SHIFT_BORDER="32", FSNET_WORKDIR=str(tmp_path / "w"),
                     FSNET_DEVICE_AUG="1").items():
        monkeypatch.setenv(k, v)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = cfg_from_file(os.path.join(repo, "configs", "kitti_wpose_files.py"))
    cfg.data.rgb_shape = (32, 64, 3)
    for aug in cfg.train_dataset.augmentation.pipeline.cfg_list:
        if aug.name.endswith("RandomWarpAffine"):
            aug.output_h, aug.output_w = 32, 64
    cfg.meta_arch.head_cfg.height, cfg.meta_arch.head_cfg.width = 32, 64
    np.random.seed(3)
    dataset = build(**cfg.train_dataset)
    stage = find_device_stage(dataset)
    loader = build_dataloader(dataset, num_workers=0, batch_size=2, collate_fn=device_augment_collate)
    backend = "tc"
    ops.set_backend("tc")
    try:
        model = build(**cfg.meta_arch).train()
        opt = build_optimizer(model, **cfg.optimizer)
        hook = build(**dict(cfg.trainer.training_hook, cuda_graph=False))
        losses = []
        for i, batch in enumerate(DevicePrefetcher(loader, device="cpu", device_transform=stage)):
            assert batch[("image", 0)].shape == (2, 3, 32, 64) and batch[("original_image", -1)].dtype == torch.float32
            assert batch["patched_mask"].dtype == torch.float64 and batch["P2"].shape == (2, 3, 4) and "frames_u8" not in batch
            assert 0.0 <= float(batch[("original_image", 1)].min()) and float(batch[("original_image", 1)].max()) <= 1.0
            losses.append(float(hook(batch, model, opt, None, None, i, 0)["loss"].detach()))
            if i == 1:
                break
    finally:
        ops.set_backend(backend)
    assert all(np.isfinite(losses)) and losses[0] != losses[1] and 0.0 < losses[0] < 1.0


def test_evaluation_hook_with_the_real_network_under_emulation(emulated, monkeypatch, tmp_path, capsys):
    """The evaluation side end to end: Eigen test reader -> BaseValidationHook -> eval-mode forward_test through the executor
    (running statistics) -> inverse-depth resize -> KittiEigenEvaluator on LiDAR ground truth exported from the miniature tree."""
    from aug_cases import eval_cfg
    from kitti_fixture import add_kitti_lidar, build_tree
    from helpers import build_model
    from fsnet_b200.networks import ops
    from vision_base.utils.builder import build
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    raw, split = build_tree(str(tmp_path))
    add_kitti_lidar(raw)
    hook = build(name="monodepth.pipeline_hooks.evaluation_hooks.base_evaluation_hooks.KittiEvaluationHook",
                 test_run_hook_cfg=dict(name="vision_base.pipeline_hooks.train_val_hooks.base_validation_hooks.BaseValidationHook"),
                 dataset_eval_cfg=dict(name="monodepth.evaluation.kitti_unsupervised_eval.KittiEigenEvaluator", data_path=raw,
                                       split_file=split, gt_saved_file=str(tmp_path / "gt.npz")),
                 num_workers=0, batch_size=4)
    ds = build(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoEigenTestDataset", raw_path=raw, split_file=split,
               augmentation=eval_cfg(size=(32, 64)))
    backend = "tc"
    ops.set_backend("tc")
    try:
        model = build_model(O.Topology(height=32, width=64)).train()
        res = hook(model, ds, None, 0, 1)
    finally:
        ops.set_backend(backend)
    assert not model.training
    assert res["error"].shape == (7,) and np.all(np.isfinite(res["error"])) and np.all(np.isfinite(res["abs_error"]))
    assert 0.0 <= res["error"][4] <= res["error"][5] <= res["error"][6] <= 1.0          # a1 <= a2 <= a3
    assert "abs_rel" in capsys.readouterr().out


def test_changing_batch_and_image_sizes_and_train_eval_transitions_under_emulation(emulated):
    """One model object through train(B=3, 32x64) -> eval(B=1) -> eval(B=4) -> train(B=2, 64x96) -> eval(64x96) -> eval(32x64):
    the executor's cached per-layer state must not depend on shapes, and the running statistics left by the training calls must
    be the ones the oracle accumulates (eval predictions to 1e-5)."""
    from helpers import build_model
    from fsnet_b200.networks import ops
    backend = "tc"
    ops.set_backend("tc")
    try:
        base = O.Topology(height=32, width=64)
        model = build_model(base)
        sd = O.make_state_dict(base)
        for mode, B, H, W in [("train", 3, 32, 64), ("eval", 1, 32, 64), ("eval", 4, 32, 64), ("train", 2, 64, 96), ("eval", 2, 64, 96),
                              ("eval", 3, 32, 64)]:
            topo = O.Topology(height=H, width=W)
            data = O.synthetic_batch(B, H, W, 5 + B + H, topo.frame_ids)
            if mode == "train":
                noise = O.tie_break_noise(B, H, W, topo.scales, 0)
                model.train()
                model.head.tie_break_noise = noise
                mine = float(model(dict(data), dict(is_training=True, epoch_num=0, global_step=0))["loss"].detach())
                ref = float(O.forward_train(sd, data, topo, noise)["loss"].detach())          # updates sd's running statistics
                assert abs(mine - ref) <= 1e-4 * abs(ref), (mode, B, H, W)
            else:
                model.eval()
                with torch.no_grad():
                    pred = model(dict(data), dict(is_training=False))["depth"]
                    want = O.forward_test(sd, data, topo)["depth"]
                assert rel(pred, want) < 1e-5, (mode, B, H, W)
    finally:
        ops.set_backend(backend)


def _conv_cases():
    from test_conv_gpu import CASES
    return CASES


@pytest.mark.parametrize("N,Cin,Cout,H,W,k,stride,pad,mode", _conv_cases())
def test_abi_level_conv_stand_ins_follow_the_contract(emulated, N, Cin, Cout, H, W, k, stride, pad, mode):
    """The two stand-ins of tests/host_emulation/conv_ref.cpp (what the executor emulation rests on) pass the very checks the
    tcgen05 kernels pass on the B200 (tests/test_conv_gpu.py: forward, fused statistics, weight gradient per tap, data gradient
    incl. the ringed variants) -- with the real planner in front and the emulated operand-plane / re-layout kernels around them."""
    from test_conv_gpu import run_conv_case
    run_conv_case("cpu", N, Cin, Cout, H, W, k, stride, pad, mode)
