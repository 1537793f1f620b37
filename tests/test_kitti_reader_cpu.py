"""KITTI raw readers (fsnet_b200/data/kitti.py under the reference's dotted names) on a miniature KITTI tree, against
golden vectors produced by the reference's own dataset classes on the same tree (tests/golden/make_golden_aug.py::run_kitti):
calibration parsing, camera-frame relative poses, static-sample filtering, left/right camera selection, sample schema, the
augmentation pipeline on real files, ground-truth depth loading."""
import os

import numpy as np
import torch

from aug_cases import nusc_train_cfg, summarize, train_cfg, val_cfg
from kitti_fixture import build_tree


def _check(prefix, got, g):
    keys = [k[len(prefix):] for k in g.files if k.startswith(prefix)]
    assert sorted(keys) == sorted(got.keys()), (sorted(keys), sorted(got.keys()))
    for k in keys:
        want, mine = g[prefix + k], got[k]
        if k.startswith("dtype/"):
            assert str(want) == str(mine), k
        elif "relative_pose" in k:
            np.testing.assert_allclose(mine, want, atol=3e-6, err_msg=k)
        else:
            np.testing.assert_allclose(mine, want, rtol=1e-6, atol=1e-6, err_msg=k)


def test_kitti_readers_match_reference(golden_dir, tmp_path):
    from vision_base.utils.builder import build
    g = np.load(os.path.join(golden_dir, "kitti_reader.npz"))
    raw, split = build_tree(str(tmp_path))
    np.random.seed(11)
    train = build(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoDataset", raw_path=raw, split_file=split,
                  frame_idxs=[0, 1, -1], is_filter_static=True, augmentation=train_cfg())
    assert len(train) == int(g["train_len"]) == 10           # two of the twelve listed samples touch the standing frame pair
    for i in (0, 3, len(train) - 1):
        _check(f"train/{i}/", summarize(train[i]), g)
    sample = train[1]
    assert sample["patched_mask"].dtype == torch.float64 and sample["P2"].shape == (3, 4) and sample[("relative_pose", 1)].shape == (4, 4)
    np.random.seed(12)
    cfg = val_cfg()
    cfg.image_keys = [("image", 0), ("image", -1), ("original_image", 0)]
    cfg.cfg_list[2].image_keys = [("image", 0), ("image", -1)]
    cfg.cfg_list[3].image_keys = [("original_image", 0)]
    cfg.gt_image_keys = []
    test = build(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoEigenTestDataset", raw_path=raw, split_file=split,
                 depth_path=raw, augmentation=cfg)
    assert len(test) == int(g["test_len"]) == 12
    for i in (0, 5):
        _check(f"test/{i}/", summarize(test[i]), g)


def test_kitti_dataset_feeds_the_dataloader(tmp_path):
    """collate_fn + build_dataloader on the reader: the batch has the schema the training hook / model consume."""
    from vision_base.utils.builder import build
    from vision_base.data.dataloader import build_dataloader
    from vision_base.data.datasets.dataset_utils import collate_fn
    raw, split = build_tree(str(tmp_path))
    ds = build(name="monodepth.data.datasets.mono_dataset.KittiDepthMonoDataset", raw_path=raw, split_file=split,
               frame_idxs=[0, 1, -1], is_filter_static=False, augmentation=train_cfg())
    loader = build_dataloader(ds, num_workers=0, batch_size=4, collate_fn=collate_fn)
    batch = next(iter(loader))
    assert batch[("image", 0)].shape == (4, 3, 48, 160) and batch[("original_image", -1)].shape == (4, 3, 48, 160)
    assert batch["P2"].shape == (4, 3, 4) and batch[("relative_pose", 1)].shape == (4, 4, 4) and batch["patched_mask"].shape == (4, 48, 160)
    assert batch["patched_mask"].dtype == torch.float64 and float(batch[("original_image", 0)].max()) <= 1.0


def test_kitti360_fisheye_reader_matches_reference(golden_dir, tmp_path):
    """KITTI360FisheyeDataset on a miniature KITTI-360 tree against the reference's reader: MEI yaml calibration -> P2 and
    calib_meta, camera-frame poses, the static / jump filter, random left / right camera, the fisheye augmentation list."""
    from vision_base.utils.builder import build
    from aug_cases import fisheye_train_cfg
    from kitti_fixture import build_kitti360_tree
    g = np.load(os.path.join(golden_dir, "kitti360_fisheye_reader.npz"))
    raw, meta, mask_path = build_kitti360_tree(str(tmp_path))
    np.random.seed(13)
    ds = build(name="monodepth.data.datasets.fisheye_dataset.KITTI360FisheyeDataset", raw_path=raw, split_file=meta,
               frame_ids=[0, 1, -1], is_filter_static=True, use_right_image=True, augmentation=fisheye_train_cfg())
    assert len(ds) == int(g["len"]) == 4                   # of 8 listed samples: two touch the standing pair, two the 5 m jump
    for i in range(len(ds)):
        s = ds[i]
        meta_d = s.pop("calib_meta")
        assert np.isclose(meta_d["mirror_parameters"]["xi"], float(g[f"{i}/xi"])) and np.isclose(meta_d["projection_parameters"]["u0"], float(g[f"{i}/u0"]))
        got = summarize(s)
        keys = [k[len(f"{i}/"):] for k in g.files if k.startswith(f"{i}/") and k[len(f"{i}/"):] not in ("xi", "u0")]
        assert sorted(keys) == sorted(got.keys())
        for k in keys:
            if k.startswith("dtype/"):
                assert str(g[f"{i}/{k}"]) == str(got[k]), k
            else:
                np.testing.assert_allclose(got[k], g[f"{i}/{k}"], rtol=1e-6, atol=3e-6, err_msg=k)
    # the validity mask (honoured from its configured path) ends up as the fp64 patched_mask
    ds2 = build(name="monodepth.data.datasets.fisheye_dataset.KITTI360FisheyeDataset", raw_path=raw, split_file=meta,
                frame_ids=[0, 1, -1], is_filter_static=False, use_right_image=False, fisheye_mask=mask_path, augmentation=fisheye_train_cfg())
    s = ds2[0]
    assert s["patched_mask"].shape == (64, 64) and 0.5 < float(s["patched_mask"].float().mean()) < 0.9
    assert s["P2"].shape == (3, 4) and float(s["P2"][0, 3]) == 0.0 and isinstance(s["calib_meta"], dict)


def test_kitti360_perspective_reader_matches_reference(golden_dir, tmp_path):
    """KITTI360MonoDataset (rectified perspective cameras) against the reference's reader on the miniature KITTI-360 tree:
    perspective.txt parsing, R_rect @ cam_to_pose extrinsics, filter, random camera, the KITTI augmentation list."""
    from vision_base.utils.builder import build
    from kitti_fixture import build_kitti360_tree
    g = np.load(os.path.join(golden_dir, "kitti360_reader.npz"))
    raw, meta, _ = build_kitti360_tree(str(tmp_path))
    np.random.seed(14)
    cfg = train_cfg()
    cfg.cfg_list[1].shift_border = 16
    ds = build(name="monodepth.data.datasets.kitti360_dataset.KITTI360MonoDataset", raw_path=raw, split_file=meta,
               frame_ids=[0, 1, -1], is_filter_static=True, use_right_image=True, augmentation=cfg)
    assert len(ds) == int(g["len"]) == 4
    for i in range(len(ds)):
        _check(f"{i}/", summarize(ds[i]), g)
    from monodepth.data.datasets.kitti360_dataset import read_extrinsic_from_sequence
    T0, T1 = read_extrinsic_from_sequence(os.path.join(raw, "calibration", "calib_cam_to_pose.txt"))
    assert T0.shape == (4, 4) and not np.allclose(T0, T1)


def test_nuscenes_json_reader_matches_reference(golden_dir, tmp_path):
    """NusceneJsonDataset against the reference's reader on a miniature JSON export (incl. the CAM_BACK ego-car rows)."""
    from vision_base.utils.builder import build
    from kitti_fixture import build_nusc_json
    g = np.load(os.path.join(golden_dir, "nusc_reader.npz"))
    path = build_nusc_json(str(tmp_path))
    np.random.seed(15)
    ds = build(name="monodepth.data.datasets.nuscene_dataset.NusceneJsonDataset", json_path=path, frame_ids=[0, 1, -1],
               augmentation=nusc_train_cfg())
    assert len(ds) == int(g["len"]) == 5
    for i in range(len(ds)):
        s = ds[i]
        meta = [s.pop("camera_type"), str(s.pop("camera_type_index")), s.pop(("filename", 0))]
        want = g[f"{i}/meta"].tolist()
        assert meta[:2] == want[:2] and meta[2].split(os.sep)[-1] == want[2].split(os.sep)[-1]
        got = summarize(s)
        keys = [k[len(f"{i}/"):] for k in g.files if k.startswith(f"{i}/") and not k.endswith("/meta")]
        assert sorted(keys) == sorted(got.keys())
        for k in keys:
            if k.startswith("dtype/"):
                assert str(g[f"{i}/{k}"]) == str(got[k]), k
            else:
                np.testing.assert_allclose(got[k], g[f"{i}/{k}"], rtol=1e-6, atol=3e-6, err_msg=k)


def test_nuscenes_config_feeds_the_dataloader(tmp_path, monkeypatch):
    """configs/nusc_wpose_files.py: both JSON exports concatenated, the recipe's pad-resize, one collated training batch and
    one single-frame evaluation sample."""
    from vision_base.utils.builder import build
    from vision_base.utils.utils import cfg_from_file
    from vision_base.data.dataloader import build_dataloader
    from vision_base.data.datasets.dataset_utils import collate_fn
    from kitti_fixture import build_nusc_json
    a = build_nusc_json(str(tmp_path / "a"), seed=3, n=3)
    b = build_nusc_json(str(tmp_path / "b"), seed=4, n=4)
    monkeypatch.setenv("FSNET_NUSC_JSON", f"{a},{b}")
    monkeypatch.setenv("FSNET_NUSC_SIZE", "64x128")
    monkeypatch.setenv("FSNET_WORKDIR", str(tmp_path / "work"))
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = cfg_from_file(os.path.join(repo, "configs", "nusc_wpose_files.py"))
    train = build(**cfg.train_dataset)
    assert len(train) == 7
    loader = build_dataloader(train, num_workers=0, batch_size=4, collate_fn=collate_fn)
    batch = next(iter(loader))
    assert batch[("image", 0)].shape == (4, 3, 64, 128) and batch[("original_image", -1)].shape == (4, 3, 64, 128)
    assert batch["P2"].shape == (4, 3, 4) and batch[("relative_pose", -1)].shape == (4, 4, 4) and batch["patched_mask"].shape == (4, 64, 128)
    assert len(batch["camera_type"]) == 4 and len(batch[("filename", 0)]) == 4
    val = build(**cfg.val_dataset)
    s = val[1]
    assert s[("image", 0)].shape == (3, 64, 128) and ("image", 1) not in s and s["camera_type"] == "CAM_BACK"
    assert cfg.meta_arch.head_cfg.depth_decoder_cfg.base_fx == 369 and cfg.meta_arch.head_cfg.overlapped_mask is False
