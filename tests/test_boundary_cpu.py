"""Host-side (no GPU) checks of the drop-in boundary: plugin names, constructor surface, state-dict
layout, config loading / overriding, the C ABI's exported symbols, and that no CPU fallback exists."""
import ctypes
import os

import numpy as np
import pytest
import torch
from easydict import EasyDict

from oracle import fsnet_oracle as O
from helpers import build_model, meta_arch_cfg

DOTTED = [
    "monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthWPose",
    "monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthMeta",
    "vision_base.networks.models.backbone.resnet.resnet",
    "monodepth.networks.models.heads.monodepth2_decoder.MonoDepth2Decoder",
    "monodepth.networks.models.heads.monodepth2_decoder.FishEyeDecoder",
    "monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoder",
    "monodepth.networks.models.heads.depth_encoder.DepthDecoder",
    "monodepth.networks.models.heads.pose_decoder.PoseDecoder",
    "vision_base.pipeline_hooks.train_val_hooks.base_training_hooks.BaseTrainingHook",
    "vision_base.pipeline_hooks.train_val_hooks.base_validation_hooks.BaseValidationHook",
    "vision_base.utils.builder.Sequential", "vision_base.utils.builder.Shuffle", "vision_base.utils.builder.Parallel",
    "vision_base.data.datasets.dataset_utils.ConcatDataset", "vision_base.data.datasets.dataset_utils.collate_fn",
    "vision_base.data.dataloader.build_dataloader", "vision_base.data.dataloader.distributed_sampler.TrainingSampler",
    "vision_base.networks.optimizers.optimizers.build_optimizer", "vision_base.networks.optimizers.schedulers.build_scheduler",
    "vision_base.networks.utils.utils.save_models", "vision_base.networks.utils.utils.load_models",
    "vision_base.networks.blocks.blocks.ConvBnReLU", "vision_base.networks.models.meta_archs.base_meta.BaseMetaArch",
    "vision_base.utils.utils.cfg_from_file", "vision_base.utils.utils.update_cfg", "vision_base.utils.logger.LossLogger",
    "vision_base.utils.timer.Timer",
] + [f"vision_base.data.augmentations.augmentations.{n}" for n in (
    "ConvertToFloat", "RandomWarpAffine", "RandomMirror", "RandomBrightness", "RandomContrast", "ConvertColor", "RandomSaturation",
    "Normalize", "ConvertToTensor", "Resize", "Copy")] + [
    "monodepth.data.datasets.mono_dataset.KittiDepthMonoDataset", "monodepth.data.datasets.mono_dataset.KittiDepthMonoEigenTestDataset",
    "monodepth.data.datasets.utils.cam_relative_pose", "monodepth.networks.utils.monodepth_utils.compute_errors",
    "monodepth.data.datasets.fisheye_dataset.KITTI360FisheyeDataset", "monodepth.data.datasets.utils.cam_relative_pose_nusc",
    "monodepth.data.datasets.kitti360_dataset.KITTI360MonoDataset",
    "monodepth.data.datasets.nuscene_dataset.NusceneJsonDataset",
    "monodepth.evaluation.kitti_unsupervised_eval.KittiEigenEvaluator", "monodepth.evaluation.kitti_unsupervised_eval.Kitti360Evaluator",
    "monodepth.pipeline_hooks.evaluation_hooks.base_evaluation_hooks.KittiEvaluationHook",
    "monodepth.pipeline_hooks.evaluation_hooks.base_evaluation_hooks.FastNuscEvaluationHook",
    "vision_base.pipeline_hooks.evaluation_hooks.base_evaluation_hooks.BaseEvaluationHook",
    "monodepth.networks.utils.monodepth_utils.generate_depth_map", "monodepth.networks.utils.monodepth_utils.project_depth_map",
    "monodepth.networks.models.meta_archs.monodepth2_model.DistillWPoseMeta",
    "monodepth.networks.models.meta_archs.teacher_model.MonoDepthInference",
    "monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoderUncertain"]


@pytest.mark.parametrize("name", DOTTED)
def test_dotted_names_resolve(name):
    from vision_base.utils.utils import find_object
    assert find_object(name) is not None


def test_find_object_error_is_module_not_found():
    from vision_base.utils.utils import find_object
    with pytest.raises(ModuleNotFoundError) as e:
        find_object("vision_base.nope.Missing")
    assert "error traces" in str(e.value)


@pytest.mark.parametrize("topo", [O.Topology(), O.Topology(posenet=True), O.Topology(depth=50), O.Topology(depth=34, n_bins=64),
                                  O.Topology(multi_channel=False, n_bins=1, scales=(0, 2)), O.Topology(distill=True)])
def test_state_dict_layout_matches_reference(topo):
    """Keys, shapes and dtypes equal the oracle's list, which loads strictly into the reference
    (tests/golden/make_golden.py ran ``load_state_dict(strict=True)`` on it)."""
    model = build_model(topo)
    mine = model.state_dict()
    ref = O.make_state_dict(topo)
    assert list(mine.keys()) == list(ref.keys())
    for k in ref:
        assert mine[k].shape == ref[k].shape and mine[k].dtype == ref[k].dtype, k
    from vision_base.networks.models.meta_archs.base_meta import BaseMetaArch
    assert isinstance(model, BaseMetaArch)
    n = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if topo == O.Topology():
        assert n == 14_364_112 or abs(n - 14.36e6) < 0.01e6        # 14.36 M
    # SyncBN conversion and re-wrapping must keep the key layout
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(build_model(topo))
    assert list(conv.state_dict().keys()) == list(ref.keys())


def test_extra_head_kwargs_become_attributes():
    cfg = meta_arch_cfg(O.Topology())
    cfg.head_cfg.pose_loss_weight = 0.5
    cfg.head_cfg.some_new_flag = "x"
    from vision_base.utils.builder import build
    m = build(**cfg)
    assert m.head.pose_loss_weight == 0.5 and m.head.some_new_flag == "x" and m.head.overlapped_mask is True


def test_wpose_with_pose_backbone_raises():
    cfg = meta_arch_cfg(O.Topology())
    cfg.pose_backbone_cfg = EasyDict(cfg.depth_backbone_cfg, num_input_images=2)
    from vision_base.utils.builder import build
    with pytest.raises(NotImplementedError):
        build(**cfg)


def test_cfg_from_file_and_update(tmp_path):
    from vision_base.utils.utils import cfg_from_file, update_cfg
    p = tmp_path / "c.py"
    p.write_text("from easydict import EasyDict as edict\ncfg = edict()\ncfg.a = 1\ncfg.b = edict(c=0, f=2)\ncfg.c = 3\n")
    cfg = cfg_from_file(str(p))
    assert isinstance(cfg, EasyDict)
    cfg = update_cfg(cfg, **{"a": 2, "b.c": 3, "d.e.f": 4, "c.g": 1})      # reference tests/test_cfg.py:18-39
    assert cfg["b"]["f"] == 2 and cfg["a"] == 2 and cfg["b"]["c"] == 3
    assert isinstance(cfg["d"]["e"], dict) and cfg["d"]["e"]["f"] == 4
    assert isinstance(cfg["c"], dict) and cfg["c"]["g"] == 1
    with pytest.raises(AssertionError):
        cfg_from_file(str(tmp_path / "c.txt"))


def test_shipped_configs_load():
    from vision_base.utils.utils import cfg_from_file
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")
    names = [f for f in os.listdir(root) if f.endswith(".py")]
    assert names
    for f in names:
        cfg = cfg_from_file(os.path.join(root, f))
        assert isinstance(cfg, EasyDict) and "meta_arch" in cfg


def test_abi_exports_every_declared_symbol():
    from fsnet_b200 import _lib
    lib = _lib.load()
    for sym in _lib.declared_symbols():
        assert hasattr(lib, sym), sym
    assert lib.fsnet_abi_version() >= 1


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors -- never route through the oracle."""
    from fsnet_b200 import _lib, functional as Fn
    t = torch.zeros(1, 3, 8, 8)
    with pytest.raises(_lib.FsnetError):
        Fn.reprojection_loss([t[:, :1]], [t[:, :1]], torch.eye(4)[None], torch.eye(4)[None], torch.zeros(1, 3, 4), t, t, t,
                             scales=[0], overlapped_mask=False)
    with pytest.raises(_lib.FsnetError):
        Fn.depth_head(torch.zeros(1, 4, 8, 8), torch.ones(4), None, False, 0.5, 100.0)
    src = open(os.path.join(os.path.dirname(_lib.__file__), "functional.py")).read()
    assert "oracle" not in src
    # the network modules are parameter containers: CPU tensors raise, nothing falls back to eager PyTorch
    from helpers import build_model
    from oracle import fsnet_oracle as O
    from fsnet_b200.networks import ops
    assert ops.COMPARATOR is None
    model = build_model(O.Topology(height=32, width=64))
    data = O.synthetic_batch(2, 32, 64, 1)
    with pytest.raises(RuntimeError, match="no CPU"):
        model(data, dict(is_training=True, epoch_num=0, global_step=0))
    with pytest.raises(RuntimeError, match="parameter container"):
        model.depth_backbone.layer1[0](torch.zeros(1, 64, 8, 8))
    with pytest.raises(TypeError):
        model.head.depth_decoder([torch.zeros(1, c, 4, 4) for c in (64, 64, 128, 256, 512)])
    pkg = os.path.dirname(_lib.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_sampler_and_collate():
    from vision_base.data.dataloader.distributed_sampler import TrainingSampler
    from vision_base.data.datasets.dataset_utils import collate_fn
    a = list(TrainingSampler(10, rank=0, world_size=2))
    b = list(TrainingSampler(10, rank=1, world_size=2))
    assert sorted(a + b) == list(range(10)) and len(a) == 5
    batch = collate_fn([{"x": torch.ones(2), ("k", 0): np.zeros(3), "s": "a", "only0": 1}, {"x": torch.ones(2), ("k", 0): np.ones(3), "s": "b"}])
    assert batch["x"].shape == (2, 2) and batch[("k", 0)].shape == (2, 3) and batch["s"] == ["a", "b"] and "only0" not in batch


def test_synthetic_dataset_schema():
    from fsnet_b200.data.synthetic import SyntheticTripletDataset, make_batch
    ds = SyntheticTripletDataset(length=4, height=32, width=64)
    s = ds[1]
    assert s[("image", 0)].shape == (3, 32, 64) and s["patched_mask"].dtype == torch.float64 and s["P2"].shape == (3, 4)
    # the product's generator and the oracle's are the same recipe (the bench compares the two arms on equal inputs)
    mine, ref = make_batch(2, 32, 64, seed=7), O.synthetic_batch(2, 32, 64, seed=7)
    for k in ref:
        assert torch.equal(mine[k], ref[k]), k
    # fisheye variant: calib_meta dicts survive the reference-style collate as a list (dataset_utils.py:24-25)
    from fsnet_b200.data.synthetic import make_fisheye_batch
    from vision_base.data.datasets.dataset_utils import collate_fn
    fe = SyntheticTripletDataset(length=4, height=32, width=32, fisheye=True)
    batch = collate_fn([fe[0], fe[1]])
    assert isinstance(batch["calib_meta"], list) and batch["calib_meta"][1]["mirror_parameters"]["xi"] > 2
    mine, ref = make_fisheye_batch(2, 32, 32, seed=7), O.synthetic_fisheye_batch(2, 32, 32, seed=7)
    for k in ref:
        if torch.is_tensor(ref[k]):
            assert torch.equal(mine[k], ref[k]), k
    assert mine["calib_meta"] == ref["calib_meta"]


def test_fused_adam_cpu_parameters_use_stock_adam():
    """FusedAdam has no CPU kernels: parameters that are not CUDA tensors take torch.optim.Adam's own step (incl. the
    clip_grad_norm_ the hook folds into step(max_norm=...)), so build_optimizer(name='adam') stays usable in CPU tests."""
    from vision_base.networks.optimizers.optimizers import build_optimizer
    torch.manual_seed(0)
    m1, m2 = torch.nn.Linear(5, 3), torch.nn.Linear(5, 3)
    m2.load_state_dict(m1.state_dict())
    o1 = build_optimizer(m1, name="adam", lr=1e-2)
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-2)
    x = torch.randn(4, 5)
    for _ in range(3):
        for m, o in ((m1, o1), (m2, o2)):
            o.zero_grad()
            (m(x) ** 2).sum().backward()
        o1.step(max_norm=0.5)
        torch.nn.utils.clip_grad_norm_(m2.parameters(), 0.5)
        o2.step()
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=1e-7)
    assert set(o1.state_dict()["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_lazy_batch_is_a_transparent_dict():
    """The graphed hook hands the model a _LazyBatch: without copy events it must behave exactly like the dict the
    reference's models index with tuple keys (getitem / get / in / items / dict())."""
    from fsnet_b200.hooks.training import _LazyBatch
    base = {("image", 0): 1, "P2": 2, "calib_meta": [dict(a=1)]}
    d = _LazyBatch(base, {})
    assert d[("image", 0)] == 1 and d.get("missing", 7) == 7 and d.get("P2") == 2 and "P2" in d and "x" not in d
    assert dict(d) == base and sorted(map(str, d.keys())) == sorted(map(str, base.keys())) and len(list(d.items())) == 3
    assert d.copy() == base


def test_device_prefetcher_is_one_batch_ahead_and_order_preserving():
    """Host logic of the upload look-ahead (the CUDA side is stream plumbing around the same queue)."""
    from fsnet_b200.data.loading import DevicePrefetcher
    pulled = []

    def source():
        for i in range(4):
            pulled.append(i)
            yield {"x": torch.full((2,), float(i)), "name": f"s{i}"}

    seen = []
    for batch in DevicePrefetcher(source(), device="cpu"):
        seen.append(int(batch["x"][0]))
        assert batch["name"] == f"s{seen[-1]}"
        assert len(pulled) == min(seen[-1] + 2, 4)            # the next batch was requested before this one was handed over
    assert seen == [0, 1, 2, 3]
    assert [int(b["x"][0]) for b in DevicePrefetcher(source(), device="cpu", depth=3)] == [0, 1, 2, 3]
    assert list(DevicePrefetcher(iter(()), device="cpu")) == []


def test_distill_loss_autograd_plumbing(monkeypatch):
    """functional.distill_loss with the kernel launch replaced by the same arithmetic in torch: checks what the Python side
    owns (argument order of the C call, unit-gradient scaling by the upstream gradient, shapes, no gradient to the teacher)."""
    from fsnet_b200 import _lib, functional as Fn
    seen = {}

    def fake_call(name, p, t, l, n, out, gp, gu, uz):
        assert name == "fsnet_distill_loss" and n.value == p.numel() and out.dtype == torch.float64 and uz is None
        seen["n"] = n.value
        err = (t - p).abs()
        sgn = torch.sign(t - p)
        inv = 1.0 / p.numel()
        if l is None:
            out += err.double().mean()
            gp.copy_(-sgn * inv)
        else:
            u = torch.sigmoid(l)
            out += (err / u + torch.log(u + 1e-5)).double().mean()
            gp.copy_(-sgn / u * inv)
            gu.copy_((1 / (u + 1e-5) - err / u ** 2) * u * (1 - u) * inv)

    monkeypatch.setattr(_lib, "call", fake_call)
    g = torch.Generator().manual_seed(0)
    for with_u in (True, False):
        p = (torch.rand(2, 1, 6, 8, generator=g) * 30 + 1).requires_grad_(True)
        t = (torch.rand(2, 1, 6, 8, generator=g) * 30 + 1).requires_grad_(True)
        l = torch.randn(2, 1, 6, 8, generator=g).requires_grad_(True) if with_u else None
        out = Fn.distill_loss(p, t, l)
        assert out.dtype == torch.float32 and out.dim() == 0 and seen["n"] == 96
        (out.double() * 0.3 + 1.0).backward()
        p2, l2 = p.detach().clone().requires_grad_(True), (None if l is None else l.detach().clone().requires_grad_(True))
        err = (t.detach() - p2).abs()
        ref = (err / torch.sigmoid(l2) + torch.log(torch.sigmoid(l2) + 1e-5)).mean() if with_u else err.mean()
        (ref * 0.3).backward()
        assert t.grad is None and p.grad.shape == p.shape
        torch.testing.assert_close(out, ref.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(p.grad, p2.grad, rtol=1e-4, atol=1e-8)
        if with_u:
            torch.testing.assert_close(l.grad, l2.grad, rtol=1e-4, atol=1e-8)


def test_teacher_export_feeds_the_distillation_config(tmp_path, monkeypatch):
    """stage-1 checkpoint -> monodepth/transform_teacher.py -> configs/kitti_distill_synthetic.py builds with that teacher:
    teacher weights equal the stage-1 network's, frozen, eval mode; the pose head is dropped."""
    import sys
    from vision_base.utils.builder import build
    from vision_base.utils.utils import cfg_from_file
    from vision_base.networks.utils.utils import save_models
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(repo, "monodepth"))
    from monodepth.transform_teacher import transform_teacher_model
    stage1 = build_model(O.Topology(posenet=True))
    ckpt, teacher_path = str(tmp_path / "stage1_latest.pth"), str(tmp_path / "teacher.pth")
    save_models(ckpt, stage1, None)
    teacher = transform_teacher_model(ckpt, teacher_path)
    assert teacher and all(k.startswith(("depth_backbone.", "depth_decoder.")) for k in teacher)
    monkeypatch.setenv("FSNET_TEACHER", teacher_path)
    monkeypatch.setenv("FSNET_WORKDIR", str(tmp_path / "work"))
    cfg = cfg_from_file(os.path.join(repo, "configs", "kitti_distill_synthetic.py"))
    model = build(**cfg.meta_arch).train()
    sd1 = stage1.state_dict()
    for k, v in model.teacher_net.state_dict().items():
        src = sd1["head." + k if k.startswith("depth_decoder.") else k]
        assert torch.equal(v, src), k
    assert not model.teacher_net.training and model.depth_backbone.training
    assert not any(p.requires_grad for p in model.teacher_net.parameters())
    assert model.head.distillation_loss_weight == 0.3 and model.head.is_uncertain_distill
    # the optimiser is built over model.parameters() as in the reference (optimizers.py:8): the frozen teacher's parameters
    # are in the group but never receive a gradient, and FusedAdam (like torch's Adam) steps only parameters that have one
    from fsnet_b200.optim import build_optimizer
    opt = build_optimizer(model, **cfg.optimizer)
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == sum(p.numel() for p in model.parameters())


def test_distill_kernel_device_code_emulated_on_the_host():
    """csrc/distill.cu's per-pixel arithmetic and loop (compiled for the CPU behind tests/host_emulation's shim) against torch
    autograd on the reference's expression (monodepth2_decoder.py:185-203)."""
    from host_emulation.emulate import run_distill
    g = torch.Generator().manual_seed(5)
    for n, with_u in ((1, True), (777, True), (1000, False)):
        p = (torch.rand(n, generator=g) * 40 + 1).requires_grad_(True)
        t = torch.rand(n, generator=g) * 40 + 1
        t[: n // 5] = p.detach()[: n // 5]
        l = (torch.randn(n, generator=g) * 2).requires_grad_(True) if with_u else None
        err = (t - p).abs()
        ref = (err / torch.sigmoid(l) + torch.log(torch.sigmoid(l) + 1e-5)).mean() if with_u else err.mean()
        ref.backward()
        val, gp, gl, u = run_distill(p.detach().numpy(), t.numpy(), None if l is None else l.detach().numpy())
        assert abs(val - float(ref.detach())) <= 1e-5 * abs(float(ref.detach())) + 1e-7
        np.testing.assert_allclose(gp, p.grad.numpy(), rtol=1e-4, atol=1e-9)
        if with_u:
            np.testing.assert_allclose(gl, l.grad.numpy(), rtol=2e-4, atol=1e-8)
            np.testing.assert_allclose(u, torch.sigmoid(l).detach().numpy(), rtol=1e-6)


def test_train_script_loader_stages(monkeypatch):
    """scripts/train.py::build_train_loader: FSNET_PREFETCH=0 gives the reference's plain DataLoader (upload inside the hook);
    the default puts the upload prefetcher in front of the hook (no device stage, default collate, for the synthetic configs)."""
    import importlib.util
    import sys
    from torch.utils.data import DataLoader
    from vision_base.utils.builder import build
    from vision_base.utils.utils import cfg_from_file
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(repo, "scripts"))
    spec = importlib.util.spec_from_file_location("fsnet_train_script", os.path.join(repo, "scripts", "train.py"))
    train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(train)
    monkeypatch.setenv("FSNET_PREFETCH", "0")
    cfg = cfg_from_file(os.path.join(repo, "configs", "kitti_wpose_synthetic.py"))
    cfg.data.num_workers, cfg.data.batch_size = 0, 2
    cfg.train_dataset.length, cfg.train_dataset.height, cfg.train_dataset.width = 6, 32, 64
    ds = build(**cfg.train_dataset)
    loader = train.build_train_loader(cfg, ds, -1, 1, torch.device("cpu"))
    assert isinstance(loader, DataLoader) and not loader.pin_memory
    batch = next(iter(loader))
    assert batch[("image", 0)].shape == (2, 3, 32, 64) and len(loader) == 3
    monkeypatch.delenv("FSNET_PREFETCH", raising=False)
    from fsnet_b200.data.loading import DevicePrefetcher
    monkeypatch.setattr(torch.utils.data.DataLoader, "__init__", (lambda orig: lambda self, *a, **k: orig(self, *a, **{**k, "pin_memory": False}))(DataLoader.__init__))
    pf = train.build_train_loader(cfg, ds, -1, 1, torch.device("cpu"))
    assert isinstance(pf, DevicePrefetcher) and pf.device_transform is None and len(list(pf)) == 3
