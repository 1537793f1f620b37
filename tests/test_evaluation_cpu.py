"""Evaluation side: LiDAR ground-truth projection, the Eigen-protocol evaluators and the
evaluation hooks -- host code, pinned against the reference's own classes run on the miniature trees
(tests/golden/evaluators.npz, written by tests/golden/make_golden_aug.py::run_eval)."""
import os

import numpy as np
import pytest
import torch

from kitti_fixture import DATE, DRIVES, add_kitti360_lidar, add_kitti_lidar, build_kitti360_tree, build_tree


def eval_predictions(n, seed=41, shape=(48, 160)):
    g = np.random.default_rng(seed)
    return [np.exp(g.uniform(np.log(3.0), np.log(60.0), size=shape)).astype(np.float32) for _ in range(n)]


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "evaluators.npz"))


def test_lidar_depth_maps_match_reference(golden, tmp_path):
    from monodepth.networks.utils.monodepth_utils import generate_depth_map
    raw, _ = build_tree(str(tmp_path))
    add_kitti_lidar(raw)
    velo = os.path.join(raw, DATE, DRIVES[0], "velodyne_points/data", "%010d.bin" % 2)
    for cam in (2, 3):
        for vd in (False, True):
            got = generate_depth_map(os.path.join(raw, DATE), velo, cam, vd).astype(np.float32)
            want = golden[f"depth_map/cam{cam}_vel{int(vd)}"]
            assert got.shape == want.shape and (want > 0).sum() > 1000
            np.testing.assert_array_equal(got, want)


def test_zbuffer_keeps_nearest_and_zeroes_negative():
    from fsnet_b200.utils.lidar import project_depth_map
    P = np.array([[10.0, 0, 5, 0], [0, 10.0, 5, 0], [0, 0, 1, 0]])
    # (forward, left, up): the camera looks along 'up' here; two points on one pixel, one behind the sensor, one out of view
    velo = np.array([[6.0, 1.0, 2.0, 0.3], [3.0, 0.5, 1.0, 0.9], [-1.0, 0.1, 1.0, 0], [2.0, 50.0, 1.0, 0], [4.0, 0.2, -1.0, 0]], dtype=np.float32)
    depth = project_depth_map(velo, P, np.array([16, 128]))
    assert depth.shape == (16, 128) and (depth > 0).sum() == 1
    assert depth[int(round(10 * 0.5 + 5)) - 1, int(round(10 * 3.0 + 5)) - 1] == 3.0          # the nearer of the two, at round(u) - 1


def test_kitti_eigen_evaluator_matches_reference(golden, tmp_path, capsys):
    from vision_base.utils.builder import build
    raw, split = build_tree(str(tmp_path))
    add_kitti_lidar(raw)
    gt_file = str(tmp_path / "gt.npz")
    cfg = dict(name="monodepth.evaluation.kitti_unsupervised_eval.KittiEigenEvaluator", data_path=raw, split_file=split, gt_saved_file=gt_file)
    ev = build(**cfg)
    n = int(golden["kitti/n"])
    assert len(ev.gt_depths) == n and os.path.isfile(gt_file)
    for i in (0, n - 1):
        np.testing.assert_array_equal(np.asarray(ev.gt_depths[i]), golden[f"kitti/gt{i}"])
    res = [ev.single_call(p, i) for i, p in enumerate(eval_predictions(n))]
    np.testing.assert_allclose([r["ratio"] for r in res], golden["kitti/ratio"], rtol=1e-6)
    np.testing.assert_allclose([r["error"] for r in res], golden["kitti/error"], rtol=1e-5)
    np.testing.assert_allclose([r["abs_error"] for r in res], golden["kitti/abs_error"], rtol=1e-5)
    capsys.readouterr()
    ev.log(None, golden["kitti/error"].mean(0), golden["kitti/abs_error"].mean(0), epoch_num=3)
    assert capsys.readouterr().out == str(golden["kitti/log"])
    again = build(**cfg)                                     # second construction reads the exported file
    np.testing.assert_allclose(again.single_call(eval_predictions(1)[0], 0)["error"], golden["kitti/reload_error"], rtol=1e-5)
    with pytest.raises(ValueError):                          # no LiDAR return inside the crop
        ev._single_loss(eval_predictions(1)[0], np.zeros((120, 400), dtype=np.float32))


def test_kitti_evaluator_on_a_directory_of_depth_pngs(golden, tmp_path, capsys):
    """__call__(result_path): 16-bit pngs (depth * 256) in sorted order against the same ground truth."""
    import cv2
    from fsnet_b200.evaluation.kitti import KittiEigenEvaluator
    raw, split = build_tree(str(tmp_path))
    add_kitti_lidar(raw)
    ev = KittiEigenEvaluator(raw, split, str(tmp_path / "gt.npz"))
    out = tmp_path / "pred"
    out.mkdir()
    preds = eval_predictions(len(ev.gt_depths))
    for i, p in enumerate(preds):
        cv2.imwrite(str(out / ("%06d.png" % i)), (p * 256).astype(np.uint16))
    res = ev(str(out), epoch_num=1)
    quant = [ev._single_loss((p * 256).astype(np.uint16).astype(np.float32) / 256.0, ev.gt_depths[i]) for i, p in enumerate(preds)]
    np.testing.assert_allclose(res["error"], np.array([q["error"] for q in quant]).mean(0), rtol=1e-6)
    np.testing.assert_allclose(res["error"], golden["kitti/error"].mean(0), rtol=2e-2)       # only the png quantisation differs
    assert "Scaled Error" in capsys.readouterr().out
    (out / "extra.png").write_bytes(b"")
    assert ev(str(out)) is None                               # count mismatch: evaluation dropped, as in the reference


def test_kitti360_evaluator_matches_reference(golden, tmp_path):
    from vision_base.utils.builder import build
    raw, meta, _ = build_kitti360_tree(str(tmp_path))
    add_kitti360_lidar(raw)
    ev = build(name="monodepth.evaluation.kitti_unsupervised_eval.Kitti360Evaluator", data_path=raw, split_file=meta,
               gt_saved_file=str(tmp_path / "gt360.npz"))
    n = int(golden["kitti360/n"])
    assert len(ev.gt_depths) == n
    np.testing.assert_array_equal(np.asarray(ev.gt_depths[0]), golden["kitti360/gt0"])
    res = [ev.single_call(p, i) for i, p in enumerate(eval_predictions(n, seed=42, shape=(32, 104)))]
    np.testing.assert_allclose([r["error"] for r in res], golden["kitti360/error"], rtol=1e-5)
    np.testing.assert_allclose([r["abs_error"] for r in res], golden["kitti360/abs_error"], rtol=1e-5)


# ----------------------------------------------------------------------------------------------------------------------
# hooks: the network is replaced by a stub (the CUDA model is exercised by the -m gpu tests); the bookkeeping is what is tested
# ----------------------------------------------------------------------------------------------------------------------
class StubValidationHook:
    """Stands in for BaseValidationHook on a machine without a GPU: same call signature, no .cuda()."""

    def __init__(self, **kwargs):
        pass

    def __call__(self, data, meta_arch, global_step=0, epoch_num=0):
        return meta_arch(data, dict(epoch_num=epoch_num, global_step=global_step, is_training=False))


class StubDepthNet(torch.nn.Module):
    """Returns the k-th synthetic prediction for the k-th sample it sees, padded like the pad-resize would."""

    def __init__(self, preds, pad=(0, 0)):
        super().__init__()
        self.preds, self.pad, self.seen = preds, pad, 0

    def forward(self, data, meta):
        assert not meta["is_training"] and not self.training
        b = data[("image", 0)].shape[0]
        out = []
        for _ in range(b):
            p = torch.from_numpy(self.preds[self.seen])
            out.append(torch.nn.functional.pad(p, (0, self.pad[1], 0, self.pad[0]), value=1e4))
            self.seen += 1
        return dict(depth=torch.stack(out)[:, None])


def test_kitti_evaluation_hook_runs_the_eigen_protocol(golden, tmp_path, capsys):
    """KittiEvaluationHook over the Eigen test reader: crop to the effective size, inverse-depth resize to the original
    frame, k-th prediction against k-th ground truth, mean over the split."""
    import cv2
    from vision_base.utils.builder import build
    from aug_cases import eval_cfg
    raw, split = build_tree(str(tmp_path))
    add_kitti_lidar(raw)
    hook = build(name="monodepth.pipeline_hooks.evaluation_hooks.base_evaluation_hooks.KittiEvaluationHook",
                 test_run_hook_cfg=dict(name="test_evaluation_cpu.StubValidationHook"),
                 dataset_eval_cfg=dict(name="monodepth.evaluation.kitti_unsupervised_eval.KittiEigenEvaluator", data_path=raw,
                                       split_file=split, gt_saved_file=str(tmp_path / "gt.npz")),
                 result_path_split="validation", num_workers=0, batch_size=2)
    ds = build(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoEigenTestDataset", raw_path=raw, split_file=split,
               augmentation=eval_cfg())
    n = len(ds)
    assert n == int(golden["kitti/n"])
    s = ds[0]
    h_eff, w_eff = (int(v) for v in s[("image_resize", "effective_size")])
    preds = eval_predictions(n, shape=(h_eff, w_eff))
    net = StubDepthNet(preds, pad=(3, 5)).train()
    res = hook(net, ds, None, 0, 2)
    assert net.seen == n and not net.training
    ev = hook.dataset_eval_func
    want = [ev.single_call(1 / cv2.resize(1 / p, (400, 120)), i) for i, p in enumerate(preds)]
    np.testing.assert_allclose(res["error"], np.array([w["error"] for w in want]).mean(0), rtol=1e-6)
    np.testing.assert_allclose(res["abs_error"], np.array([w["abs_error"] for w in want]).mean(0), rtol=1e-6)
    assert "Epoch 2" in capsys.readouterr().out


class _CamEvaluator:
    """nuScenes-style evaluator stub: ground truth by file name; one frame has no usable points."""

    def __init__(self):
        self.logged = []

    def single_call(self, depth, filename):
        if filename.endswith("n003_frame0.png"):
            raise ValueError
        v = float(depth.mean())
        return dict(error=np.full(7, v), abs_error=np.full(7, 2 * v), shape=depth.shape)

    def log(self, writer, cam, mean_errors, mean_abs_errors, global_step=0, epoch_num=0):
        self.logged.append((cam, mean_errors.copy(), mean_abs_errors.copy()))


def test_fast_nusc_evaluation_hook_groups_by_camera(tmp_path):
    from vision_base.utils.builder import build
    from aug_cases import eval_cfg
    from kitti_fixture import build_nusc_json
    path = build_nusc_json(str(tmp_path), n=6, h0=96)
    cfg = eval_cfg(size=(64, 128), preserve_aspect_ratio=True)       # 96x160 frames -> 64x107 padded to 128
    ds = build(name="monodepth.data.datasets.nuscene_dataset.NusceneJsonDataset", json_path=path, image_keys=["frame0"], frame_ids=[0],
               augmentation=cfg)
    ds.json_dict["samples"] = [s for s in ds.json_dict["samples"] if s["camera_type"] != "CAM_BACK"]      # equal frame sizes: batchable
    hook = build(name="monodepth.pipeline_hooks.evaluation_hooks.base_evaluation_hooks.FastNuscEvaluationHook",
                 test_run_hook_cfg=dict(name="test_evaluation_cpu.StubValidationHook"), num_workers=0, batch_size=3)
    hook.dataset_eval_func = _CamEvaluator()
    s = ds[0]
    h_eff, w_eff = (int(v) for v in s[("image_resize", "effective_size")])
    preds = eval_predictions(len(ds), shape=(h_eff, w_eff))
    assert (h_eff, w_eff) == (64, 107)
    with pytest.warns(UserWarning, match="no usable points"):
        res = hook(StubDepthNet(preds, pad=(0, 128 - w_eff)), ds, None, 0, 0)
    front = float(__import__("cv2").resize(preds[0], (160, 96)).mean())                # sample 3 (the other CAM_FRONT frame) was skipped
    cams = [c for c, _, _ in hook.dataset_eval_func.logged]
    assert cams == ["CAM_FRONT", "CAM_FRONT_LEFT", "all mean"]
    per_cam = {c: e for c, e, _ in hook.dataset_eval_func.logged}
    np.testing.assert_allclose(per_cam["CAM_FRONT"], front, rtol=1e-6)
    np.testing.assert_allclose(res["error"], (per_cam["CAM_FRONT"] + per_cam["CAM_FRONT_LEFT"]) / 2)
    np.testing.assert_allclose(res["abs_error"], 2 * res["error"])


def test_base_evaluation_hook_steps_every_sample():
    from vision_base.utils.builder import build

    hook = build(name="vision_base.pipeline_hooks.evaluation_hooks.base_evaluation_hooks.BaseEvaluationHook",
                 test_run_hook_cfg=dict(name="test_evaluation_cpu.StubValidationHook"),
                 dataset_eval_cfg=dict(name="test_evaluation_cpu.RecordingEvaluator"), extra=5)
    assert hook.extra == 5 and hook.result_path_split == "validation"
    ds = [{("image", 0): torch.zeros(3, 4, 6), "k": i} for i in range(3)]
    hook(StubDepthNet([np.ones((4, 6), dtype=np.float32)] * 3), ds, None, 7, 1)
    ev = hook.dataset_eval
    assert ev.events == ["reset", 0, 1, 2, ("final", 7, 1)]
    hook.result_path_split = "test"
    hook(StubDepthNet([np.ones((4, 6), dtype=np.float32)] * 3), ds, None, 7, 1)
    assert ev.events == ["reset", 0, 1, 2]


class RecordingEvaluator:
    def __init__(self):
        self.events = []

    def reset(self):
        self.events = ["reset"]

    def step(self, index, output, data):
        assert output["depth"].shape == (1, 1, 4, 6) and data["k"] == index
        self.events.append(index)

    def __call__(self, writer, global_step, epoch_num):
        self.events.append(("final", global_step, epoch_num))
