"""Golden vectors for the host-side augmentation pipeline, produced by RUNNING THE REFERENCE's classes
(vision_base/data/augmentations/augmentations.py) through its own builder with the train / val augmentation lists of
configs/kitti_wpose_example.  Build container only:   python tests/golden/make_golden_aug.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") not in (REPO, HERE)] + [REPO]

import numpy as np  # noqa: E402

np.int = int                       # the reference's Resize uses the alias numpy 2 removed (SURVEY.md 8(c) caveat iv)
from easydict import EasyDict as edict  # noqa: E402
from vision_base.utils.builder import build  # noqa: E402  (reference)
import vision_base  # noqa: E402

assert vision_base.__path__[0].startswith(REF)
sys.path.insert(0, os.path.join(REPO, "tests"))
from aug_cases import raw_sample, train_cfg, val_cfg, summarize, nusc_train_cfg  # noqa: E402


def run(name, cfg, seed):
    np.random.seed(seed)
    pipe = build(**cfg)
    outs = {}
    for i in range(3):                         # three consecutive samples: exercises the generators' state
        out = pipe(raw_sample(100 + i))
        for k, v in summarize(out).items():
            outs[f"{i}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **outs)
    print(name, len(outs), "entries", os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    run("aug_train", train_cfg(), 7)
    run("aug_val", val_cfg(), 8)


def run_metrics():
    """compute_errors / compute_depth_errors / disp<->depth of the reference and KittiEigenEvaluator._single_loss."""
    import torch
    from monodepth.networks.utils import monodepth_utils as U
    from monodepth.evaluation.kitti_unsupervised_eval import KittiEigenEvaluator
    g = np.random.default_rng(21)
    gt = g.uniform(1.0, 70.0, size=5000)
    pred = gt * g.uniform(0.7, 1.4, size=5000)
    out = {"errors": np.array(U.compute_errors(gt, pred)),
           "torch_errors": np.array([float(v) for v in U.compute_depth_errors(torch.tensor(gt), torch.tensor(pred))]),
           "depth": U.disp_to_depth(np.linspace(0, 1, 11), 0.1, 100.0)[1], "disp": U.depth_to_disp(np.linspace(0.5, 90, 11), 0.1, 100.0)}
    gt_map = np.where(g.uniform(size=(94, 310)) < 0.2, g.uniform(0.5, 90.0, size=(94, 310)), 0.0).astype(np.float32)
    pred_map = g.uniform(2.0, 60.0, size=(48, 160)).astype(np.float32)
    ev = KittiEigenEvaluator.__new__(KittiEigenEvaluator)          # the constructor wants KITTI on disk; the protocol does not
    res = ev._single_loss(pred_map.copy(), gt_map.copy())
    out["eigen_ratio"] = np.array(res["ratio"]); out["eigen_error"] = np.array(res["error"]); out["eigen_abs_error"] = np.array(res["abs_error"])
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("metrics", out["errors"])


if __name__ == "__main__":
    run_metrics()


def run_kitti():
    """The reference's KITTI readers on the miniature tree of tests/kitti_fixture.py."""
    import tempfile
    from kitti_fixture import build_tree
    with tempfile.TemporaryDirectory() as root:
        raw, split = build_tree(root)
        out = {}
        np.random.seed(11)
        train = build(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoDataset", raw_path=raw, split_file=split,
                      frame_idxs=[0, 1, -1], is_filter_static=True, augmentation=train_cfg())
        out["train_len"] = np.array(len(train))
        for i in (0, 3, len(train) - 1):
            for k, v in summarize(train[i]).items():
                out[f"train/{i}/{k}"] = v
        np.random.seed(12)
        cfg = val_cfg()
        cfg.image_keys = [("image", 0), ("image", -1), ("original_image", 0)]
        cfg.cfg_list[2].image_keys = [("image", 0), ("image", -1)]
        cfg.cfg_list[3].image_keys = [("original_image", 0)]
        cfg.gt_image_keys = []
        test = build(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoEigenTestDataset", raw_path=raw, split_file=split,
                     depth_path=raw, augmentation=cfg)
        out["test_len"] = np.array(len(test))
        for i in (0, 5):
            for k, v in summarize(test[i]).items():
                out[f"test/{i}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "kitti_reader.npz"), **out)
    print("kitti", int(out["train_len"]), int(out["test_len"]), len(out), "entries")


if __name__ == "__main__":
    run_kitti()


def run_fisheye():
    """The reference's KITTI-360 fisheye reader on the miniature tree (its hard-coded mask path is not used: no mask key)."""
    import tempfile
    from kitti_fixture import build_kitti360_tree
    from aug_cases import fisheye_train_cfg
    with tempfile.TemporaryDirectory() as root:
        raw, meta, _ = build_kitti360_tree(root)
        np.random.seed(13)
        ds = build(name="monodepth.data.datasets.fisheye_dataset.KITTI360FisheyeDataset", raw_path=raw, split_file=meta,
                   frame_ids=[0, 1, -1], is_filter_static=True, use_right_image=True, augmentation=fisheye_train_cfg())
        out = {"len": np.array(len(ds))}
        for i in range(len(ds)):
            s = ds[i]
            meta_d = s.pop("calib_meta")
            out[f"{i}/xi"] = np.array(meta_d["mirror_parameters"]["xi"])
            out[f"{i}/u0"] = np.array(meta_d["projection_parameters"]["u0"])
            for k, v in summarize(s).items():
                out[f"{i}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "kitti360_fisheye_reader.npz"), **out)
    print("fisheye reader", int(out["len"]), len(out), "entries")


if __name__ == "__main__":
    run_fisheye()


def run_kitti360():
    """The reference's KITTI-360 perspective reader on the miniature tree."""
    import tempfile
    from kitti_fixture import build_kitti360_tree
    with tempfile.TemporaryDirectory() as root:
        raw, meta, _ = build_kitti360_tree(root)
        np.random.seed(14)
        cfg = train_cfg()
        cfg.cfg_list[1].shift_border = 16
        ds = build(name="monodepth.data.datasets.kitti360_dataset.KITTI360MonoDataset", raw_path=raw, split_file=meta,
                   frame_ids=[0, 1, -1], is_filter_static=True, use_right_image=True, augmentation=cfg)
        out = {"len": np.array(len(ds))}
        for i in range(len(ds)):
            for k, v in summarize(ds[i]).items():
                out[f"{i}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "kitti360_reader.npz"), **out)
    print("kitti360 reader", int(out["len"]), len(out), "entries")


if __name__ == "__main__":
    run_kitti360()


def run_nusc():
    """The reference's NusceneJsonDataset on the miniature JSON export."""
    import tempfile
    from kitti_fixture import build_nusc_json
    with tempfile.TemporaryDirectory() as root:
        path = build_nusc_json(root)
        np.random.seed(15)
        ds = build(name="monodepth.data.datasets.nuscene_dataset.NusceneJsonDataset", json_path=path, frame_ids=[0, 1, -1],
                   augmentation=nusc_train_cfg())
        out = {"len": np.array(len(ds))}
        for i in range(len(ds)):
            s = ds[i]
            out[f"{i}/meta"] = np.array([s.pop("camera_type"), str(s.pop("camera_type_index")), s.pop(("filename", 0))])
            for k, v in summarize(s).items():
                out[f"{i}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "nusc_reader.npz"), **out)
    print("nusc reader", int(out["len"]), len(out), "entries")


if __name__ == "__main__":
    run_nusc()


def eval_predictions(n, seed=41, shape=(48, 160)):
    """The synthetic 'network outputs' both sides evaluate: smooth positive depth maps."""
    g = np.random.default_rng(seed)
    return [np.exp(g.uniform(np.log(3.0), np.log(60.0), size=shape)).astype(np.float32) for _ in range(n)]


def run_eval():
    """The reference's LiDAR projection and evaluators on the miniature trees."""
    import io
    import tempfile
    from contextlib import redirect_stdout
    from kitti_fixture import build_tree, build_kitti360_tree, add_kitti_lidar, add_kitti360_lidar, DATE, DRIVES
    from monodepth.networks.utils import monodepth_utils as U
    from monodepth.evaluation.kitti_unsupervised_eval import KittiEigenEvaluator, Kitti360Evaluator
    out = {}
    with tempfile.TemporaryDirectory() as root:
        raw, split = build_tree(root)
        add_kitti_lidar(raw)
        velo = os.path.join(raw, DATE, DRIVES[0], "velodyne_points/data", "%010d.bin" % 2)
        for cam in (2, 3):
            for vd in (False, True):
                out[f"depth_map/cam{cam}_vel{int(vd)}"] = U.generate_depth_map(os.path.join(raw, DATE), velo, cam, vd).astype(np.float32)
        with redirect_stdout(io.StringIO()):
            ev = KittiEigenEvaluator(raw, split, os.path.join(root, "gt.npz"))
        n = len(ev.gt_depths)
        out["kitti/n"] = np.array(n)
        for i in (0, n - 1):
            out[f"kitti/gt{i}"] = np.asarray(ev.gt_depths[i])
        res = [ev.single_call(p, i) for i, p in enumerate(eval_predictions(n))]
        out["kitti/ratio"] = np.array([r["ratio"] for r in res])
        out["kitti/error"] = np.array([r["error"] for r in res])
        out["kitti/abs_error"] = np.array([r["abs_error"] for r in res])
        buf = io.StringIO()
        with redirect_stdout(buf):
            ev.log(None, out["kitti/error"].mean(0), out["kitti/abs_error"].mean(0), epoch_num=3)
        out["kitti/log"] = np.array(buf.getvalue())
        ev2 = KittiEigenEvaluator(raw, split, os.path.join(root, "gt.npz"))           # second construction: loads the saved file
        out["kitti/reload_error"] = np.array(ev2.single_call(eval_predictions(1)[0], 0)["error"])
    with tempfile.TemporaryDirectory() as root:
        raw, meta, _ = build_kitti360_tree(root)
        add_kitti360_lidar(raw)
        with redirect_stdout(io.StringIO()):
            ev = Kitti360Evaluator(raw, meta, os.path.join(root, "gt360.npz"))
        n = len(ev.gt_depths)
        out["kitti360/n"] = np.array(n)
        out["kitti360/gt0"] = np.asarray(ev.gt_depths[0])
        res = [ev.single_call(p, i) for i, p in enumerate(eval_predictions(n, seed=42, shape=(32, 104)))]
        out["kitti360/error"] = np.array([r["error"] for r in res])
        out["kitti360/abs_error"] = np.array([r["abs_error"] for r in res])
    np.savez_compressed(os.path.join(HERE, "evaluators.npz"), **out)
    print("evaluators", {k: v.shape for k, v in out.items()}, os.path.getsize(os.path.join(HERE, "evaluators.npz")) // 1024, "KiB")


if __name__ == "__main__":
    run_eval()
