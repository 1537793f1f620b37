"""GPU input pipeline, host half + oracle, on CPU:
 * oracle/augment_oracle.py restates OpenCV's warp / HSV arithmetic: pinned against cv2 itself;
 * DeviceAugmentation draws the reference pipeline's random parameters from the same config list: its plan, fed to the oracle,
   must reproduce the golden vectors of the reference's own CPU pipeline (tests/golden/aug_train.npz)."""
import os

import cv2
import numpy as np
import pytest
import torch

from aug_cases import OUT_H, OUT_W, raw_sample, summarize, train_cfg
from oracle import augment_oracle as AO


def test_warp_oracle_is_bit_exact_with_cv2():
    g = np.random.default_rng(0)
    img = g.integers(0, 256, size=(120, 400, 3)).astype(np.uint8)
    mask = (g.uniform(size=(120, 400)) < 0.7).astype(np.float32)
    for s, tx, ty in ((0.8, -30.2, 11.7), (1.37, -100.5, -40.25), (0.61, 3.3, 5.5), (2.0, -200.0, -60.0)):
        M = np.array([[s, 0, tx], [0, s, ty]], dtype=np.float32)
        Minv = AO.invert_affine(M)
        ref = cv2.warpAffine(img.astype(np.float32), M, (160, 48), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        np.testing.assert_array_equal(AO.warp_linear(img, Minv, 48, 160), ref)
        np.testing.assert_array_equal(AO.warp_linear(img, Minv, 48, 160, mirror=True), ref[:, ::-1])
        refn = cv2.warpAffine(mask, M, (160, 48), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT)
        np.testing.assert_array_equal(AO.warp_nearest(mask, Minv, 48, 160), refn)
        # a zero-padded source with the true size given behaves like the unpadded one
        pad = np.pad(img, ((0, 9), (0, 17), (0, 0)), constant_values=255)
        np.testing.assert_array_equal(AO.warp_linear(pad, Minv, 48, 160, h0=120, w0=400), ref)


def test_hsv_oracle_matches_cv2():
    g = np.random.default_rng(1)
    img = g.uniform(-20, 290, size=(64, 96, 3)).astype(np.float32)
    img[:8] = g.integers(0, 256, size=(8, 96, 3)).astype(np.float32)
    img[8:10, :, 1] = img[8:10, :, 0]
    img[10:12] = img[10:12, :, :1]                        # grey pixels: S = 0
    hsv_ref = cv2.cvtColor(img, cv2.COLOR_RGB2HSV)
    np.testing.assert_allclose(AO.rgb2hsv(img), hsv_ref, rtol=0, atol=4e-5)
    hsv_ref[..., 1] *= np.float32(1.23)
    np.testing.assert_allclose(AO.hsv2rgb(hsv_ref), cv2.cvtColor(hsv_ref, cv2.COLOR_HSV2RGB), rtol=0, atol=4e-5)


def device_cfg():
    from easydict import EasyDict as edict
    return edict(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=train_cfg())


def test_plan_plus_oracle_reproduces_the_reference_pipeline(golden_dir):
    from vision_base.utils.builder import build
    g = np.load(os.path.join(golden_dir, "aug_train.npz"))
    np.random.seed(7)                                     # the seed of the golden run: same objects, same draws
    aug = build(**device_cfg())
    assert aug.frames == [0, 1, -1] and (aug.output_h, aug.output_w) == (OUT_H, OUT_W)
    for i in range(3):
        raw = raw_sample(100 + i)
        frames_in = [raw[("image", f)].copy() for f in aug.frames]
        mask_in = raw["patched_mask"].copy()
        out = aug(raw)
        plan = out["aug_plan"]
        assert out["frames_u8"].dtype == np.uint8 and out["frames_u8"].shape == (3, 120, 400, 3) and ("image", 0) not in out
        np.testing.assert_array_equal(out["frames_u8"], np.stack(frames_in))
        image, original, mask = AO.apply_plan(out["frames_u8"], out["mask_u8"], plan, OUT_H, OUT_W, aug.mean, aug.std)
        sample = {("image", f): torch.from_numpy(image[k]) for k, f in enumerate(aug.frames)}
        sample.update({("original_image", f): torch.from_numpy(original[k]) for k, f in enumerate(aug.frames)})
        sample.update({k: v for k, v in out.items() if k not in ("frames_u8", "mask_u8", "aug_plan", "mask_dtype")})
        sample["patched_mask"] = torch.from_numpy(mask)
        mine = summarize(sample)
        keys = [k[len(f"{i}/"):] for k in g.files if k.startswith(f"{i}/")]
        assert sorted(keys) == sorted(mine.keys())
        for k in keys:
            want, got = g[f"{i}/{k}"], mine[k]
            if k.startswith("dtype/"):
                assert str(want) == str(got), k
            elif "patched_mask" in k:
                continue                                  # see below
            elif k.startswith("full/relative_pose"):
                np.testing.assert_allclose(got, want, atol=2e-6)
            elif k.startswith("sum/"):
                np.testing.assert_allclose(got, want, rtol=2e-5, err_msg=k)
            else:
                np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-5, err_msg=k)
        # the mask follows cv2's documented nearest-neighbour warp (what it does for every depth but CV_64F in 4.13)
        s = 1.0 / plan[0]
        M = np.array([[s, 0, -plan[2] * s], [0, s, -plan[5] * s]], dtype=np.float32)
        ref_mask = cv2.warpAffine(mask_in.astype(np.float32), M, (OUT_W, OUT_H), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT)
        if plan[6]:
            ref_mask = ref_mask[:, ::-1]
        assert mask.dtype == np.float64 and (mask != ref_mask).mean() < 0.01


def test_collate_pads_ragged_sources():
    from fsnet_b200.data.device_augment import device_augment_collate
    from vision_base.utils.builder import build
    np.random.seed(3)
    aug = build(**device_cfg())
    a, b = aug(raw_sample(1, h=120, w=400)), aug(raw_sample(2, h=112, w=392))
    batch = device_augment_collate([a, b])
    assert batch["frames_u8"].shape == (2, 3, 120, 400, 3) and batch["frames_u8"].dtype == torch.uint8
    assert batch["mask_u8"].shape == (2, 120, 400) and batch["aug_plan"].shape == (2, 16) and batch["aug_plan"].dtype == torch.float64
    assert batch["aug_plan"][1, 13:15].tolist() == [112.0, 392.0] and int(batch["frames_u8"][1, :, 112:].sum()) == 0
    assert batch["P2"].shape == (2, 3, 4) and batch[("relative_pose", 1)].shape == (2, 4, 4)


def test_unsupported_lists_are_rejected():
    from easydict import EasyDict as edict
    from vision_base.utils.builder import build
    from aug_cases import val_cfg, AUG
    bad = train_cfg()
    bad.cfg_list.insert(3, edict(name=f"{AUG}.Resize", size=(48, 160)))                            # a second geometric step
    with pytest.raises(NotImplementedError):
        build(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=bad)
    with pytest.raises(NotImplementedError):
        build(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=edict(name="vision_base.utils.builder.Shuffle", cfg_list=[]))


def _rebuild(sample, aug):
    """plan + oracle -> the sample the reference's CPU list would have produced."""
    image, original, mask = AO.apply_plan(sample["frames_u8"], sample.get("mask_u8"), sample["aug_plan"], aug.output_h, aug.output_w,
                                          aug.mean, aug.std)
    out = {k: v for k, v in sample.items() if k not in ("frames_u8", "mask_u8", "aug_plan", "mask_dtype")}
    for k, f in enumerate(aug.frames):
        out[("image", f)] = torch.from_numpy(image[k])
        out[("original_image", f)] = torch.from_numpy(original[k])
    if mask is not None:
        out["patched_mask"] = torch.from_numpy(mask)
    return out


def _compare(got, g, i, skip=()):
    keys = [k[len(f"{i}/"):] for k in g.files if k.startswith(f"{i}/") and k[len(f"{i}/"):] not in skip]
    assert sorted(keys) == sorted(got.keys())
    for k in keys:
        if k.startswith("dtype/"):
            assert str(g[f"{i}/{k}"]) == str(got[k]), k
        elif k.startswith("sum/"):
            np.testing.assert_allclose(got[k], g[f"{i}/{k}"], rtol=2e-5, err_msg=k)
        else:
            np.testing.assert_allclose(got[k], g[f"{i}/{k}"], rtol=1e-5, atol=2e-5, err_msg=k)


def test_nuscenes_list_on_the_device_matches_the_reference_reader(golden_dir, tmp_path):
    """The nuScenes recipe (Resize + pad, colour, mirror): reader + DeviceAugmentation + oracle against the golden produced by the
    reference's reader with its CPU list (tests/golden/nusc_reader.npz) -- incl. the tall rear-camera frames (ragged sizes)."""
    from easydict import EasyDict as edict
    from vision_base.utils.builder import build
    from aug_cases import nusc_train_cfg
    from kitti_fixture import build_nusc_json
    g = np.load(os.path.join(golden_dir, "nusc_reader.npz"))
    path = build_nusc_json(str(tmp_path))
    np.random.seed(15)
    ds = build(name="monodepth.data.datasets.nuscene_dataset.NusceneJsonDataset", json_path=path, frame_ids=[0, 1, -1],
               augmentation=edict(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=nusc_train_cfg()))
    aug = ds.transform
    assert not aug.is_warp and (aug.output_h, aug.output_w) == (64, 128)
    for i in range(len(ds)):
        s = ds[i]
        assert s["aug_plan"][15] == 1 and s["frames_u8"].dtype == np.uint8
        s = _rebuild(s, aug)
        for k in ("camera_type", "camera_type_index", ("filename", 0)):
            s.pop(k)
        _compare(summarize(s), g, i, skip=("meta",))


def test_fisheye_list_on_the_device_matches_the_reference_reader(golden_dir, tmp_path):
    """The KITTI-360 fisheye recipe (Resize, mirror, Copy -> original_image, colour)."""
    from easydict import EasyDict as edict
    from vision_base.utils.builder import build
    from aug_cases import fisheye_train_cfg
    from kitti_fixture import build_kitti360_tree
    g = np.load(os.path.join(golden_dir, "kitti360_fisheye_reader.npz"))
    raw, meta, mask_path = build_kitti360_tree(str(tmp_path))
    np.random.seed(13)
    ds = build(name="monodepth.data.datasets.fisheye_dataset.KITTI360FisheyeDataset", raw_path=raw, split_file=meta,
               frame_ids=[0, 1, -1], is_filter_static=True, use_right_image=True,
               augmentation=edict(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=fisheye_train_cfg()))
    for i in range(len(ds)):
        s = _rebuild(ds[i], ds.transform)
        s.pop("calib_meta")
        _compare(summarize(s), g, i, skip=("xi", "u0"))


def test_stage_call_marshalling_and_prefetcher_hook(monkeypatch):
    """DeviceAugmentStage with the kernel launch replaced by the oracle: argument order / shapes of the C call, output keys
    and layouts, and its place inside DevicePrefetcher.  (The kernel itself: tests/test_callers_gpu.py.)"""
    from fsnet_b200 import _lib
    from fsnet_b200.data.device_augment import DeviceAugmentStage, device_augment_collate
    from fsnet_b200.data.loading import DevicePrefetcher
    from vision_base.utils.builder import build

    def fake_call(name, frames, mask, plan, B, F, H0, W0, H, W, mean_std, image, original, mask_out):
        assert name == "fsnet_augment_frames" and frames.shape == (B, F, H0, W0, 3) and plan.shape == (B, 16)
        ms = mean_std.numpy()
        for b in range(B):
            im, orig, m = AO.apply_plan(frames[b].numpy(), mask[b].numpy(), plan[b].numpy(), H, W, ms[:3], ms[3:])
            image[:, b] = torch.from_numpy(im)
            original[:, b] = torch.from_numpy(orig)
            mask_out[b] = torch.from_numpy(m)

    monkeypatch.setattr(_lib, "call", fake_call)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))       # the stage refuses host batches
    np.random.seed(7)
    aug = build(**device_cfg())
    stage = DeviceAugmentStage(aug)
    loader = [device_augment_collate([aug(raw_sample(100 + i)) for i in range(2)]) for _ in range(2)]
    batches = list(DevicePrefetcher(loader, device="cpu", device_transform=stage))
    assert len(batches) == 2
    b0 = batches[0]
    assert "frames_u8" not in b0 and "aug_plan" not in b0 and "mask_dtype" not in b0
    for f in (0, 1, -1):
        assert b0[("image", f)].shape == (2, 3, OUT_H, OUT_W) and b0[("original_image", f)].dtype == torch.float32
    assert b0["patched_mask"].shape == (2, OUT_H, OUT_W) and b0["patched_mask"].dtype == torch.float64
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aug_train.npz"))
    np.testing.assert_allclose(b0[("image", 0)][0].numpy(), g["0/full/image_0"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(b0[("original_image", 1)][1].numpy(), g["1/full/original_image_1"], rtol=1e-5, atol=2e-5)


def test_kitti_reader_with_device_augmentation(tmp_path, monkeypatch):
    """configs/kitti_wpose_files.py with FSNET_DEVICE_AUG=1: the reader hands out uint8 frames + plans, the stage is found through
    the ConcatDataset, and the drawn geometry equals what the host pipeline applies under the same seed."""
    from kitti_fixture import build_tree
    from fsnet_b200.data.device_augment import DeviceAugmentStage, device_augment_collate, find_device_stage
    from vision_base.utils.builder import build
    from vision_base.utils.utils import cfg_from_file
    raw, split = build_tree(str(tmp_path))
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for k, v in dict(FSNET_KITTI_PATH=raw, FSNET_KITTI_SPLIT=split, FSNET_SHIFT_BORDER="32", FSNET_WORKDIR=str(tmp_path / "w")).items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("FSNET_DEVICE_AUG", "1")
    np.random.seed(11)
    dev_ds = build(**cfg_from_file(os.path.join(repo, "configs", "kitti_wpose_files.py")).train_dataset)
    monkeypatch.setenv("FSNET_DEVICE_AUG", "0")
    np.random.seed(11)
    host_ds = build(**cfg_from_file(os.path.join(repo, "configs", "kitti_wpose_files.py")).train_dataset)
    assert isinstance(find_device_stage(dev_ds), DeviceAugmentStage) and find_device_stage(host_ds) is None
    np.random.seed(5)
    a = dev_ds[0]
    np.random.seed(5)
    b = host_ds[0]
    assert a["frames_u8"].dtype == np.uint8 and a["frames_u8"].shape[0] == 3 and ("image", 0) not in a
    torch.testing.assert_close(a["P2"], b["P2"])
    np.testing.assert_allclose(a[("relative_pose", 1)], b[("relative_pose", 1)], atol=1e-6)
    aug = dev_ds.children[0].transform
    image, original, mask = AO.apply_plan(a["frames_u8"], a["mask_u8"], a["aug_plan"], aug.output_h, aug.output_w, aug.mean, aug.std)
    np.testing.assert_allclose(image[0], b[("image", 0)].numpy(), rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(original[2], b[("original_image", -1)].numpy(), rtol=1e-5, atol=2e-5)
    batch = device_augment_collate([dev_ds[i] for i in range(3)])
    assert batch["frames_u8"].shape[:2] == (3, 3) and batch["aug_plan"].shape == (3, 16)


@pytest.mark.parametrize("config", ["kitti_wpose_files.py", "nusc_wpose_files.py", "kitti360_fisheye_files.py"])
def test_file_configs_accept_device_augmentation(config, monkeypatch, tmp_path):
    """FSNET_DEVICE_AUG=1 wraps the train list of every file-based config, and the wrapped list is one DeviceAugmentation accepts."""
    from fsnet_b200.data.device_augment import DeviceAugmentation
    from vision_base.utils.builder import build
    from vision_base.utils.utils import cfg_from_file
    monkeypatch.setenv("FSNET_DEVICE_AUG", "1")
    monkeypatch.setenv("FSNET_WORKDIR", str(tmp_path))
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = cfg_from_file(os.path.join(repo, "configs", config))
    aug = build(**cfg.train_dataset.augmentation)
    assert isinstance(aug, DeviceAugmentation) and aug.frames == [0, 1, -1]
    assert (aug.output_h, aug.output_w) == tuple(cfg.data.rgb_shape[:2])


def test_kernel_device_code_emulated_on_the_host_matches_oracle(tmp_path):
    """The device code of csrc/augment.cu, compiled for the CPU behind a shim (tests/host_emulation), against the oracle: all
    three shipped lists, mirrored and not, every colour-op order that gets drawn, ragged zero-padded sources, the HSV-only case.
    Checks the kernel's index arithmetic and branches in the CPU suite; the real launch is checked by the -m gpu test."""
    from easydict import EasyDict as edict
    from aug_cases import fisheye_train_cfg, nusc_train_cfg
    from host_emulation.emulate import run_augment
    from kitti_fixture import build_kitti360_tree, build_nusc_json
    from fsnet_b200.data.device_augment import device_augment_collate
    from vision_base.utils.builder import build

    def check(samples, aug, exact_original):
        host = device_augment_collate(samples)
        frames, mask, plan = host["frames_u8"].numpy(), host["mask_u8"].numpy(), host["aug_plan"].numpy()
        image, original, mask_out = run_augment(frames, mask, plan, aug.output_h, aug.output_w, np.concatenate([aug.mean, aug.std]))
        for b in range(len(samples)):
            want = AO.apply_plan(frames[b], mask[b], plan[b], aug.output_h, aug.output_w, aug.mean, aug.std)
            if exact_original:
                np.testing.assert_array_equal(original[:, b], want[1])
            else:
                np.testing.assert_allclose(original[:, b], want[1], rtol=0, atol=1e-6)
            np.testing.assert_allclose(image[:, b], want[0], rtol=1e-5, atol=2e-5)
            np.testing.assert_array_equal(mask_out[b], want[2])
        return plan

    np.random.seed(7)
    aug = build(**device_cfg())
    samples = [aug(raw_sample(100 + i, h=120 - 8 * (i % 2), w=400 - 8 * (i % 3))) for i in range(8)]
    samples[3]["aug_plan"][10:13] = np.where(samples[3]["aug_plan"][7:10] == 3, np.nan, samples[3]["aug_plan"][10:13])
    plan = check(samples, aug, exact_original=True)
    assert {0.0, 1.0} == set(plan[:, 6].tolist()) and len({tuple(p[7:10]) for p in plan}) >= 3      # mirrored / not, several orders

    np.random.seed(15)
    nusc = build(name="monodepth.data.datasets.nuscene_dataset.NusceneJsonDataset", json_path=build_nusc_json(str(tmp_path / "n")),
                 frame_ids=[0, 1, -1], augmentation=edict(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=nusc_train_cfg()))
    check([nusc[i] for i in range(5)], nusc.transform, exact_original=False)
    raw, meta, mask_path = build_kitti360_tree(str(tmp_path / "k"))
    fish = build(name="monodepth.data.datasets.fisheye_dataset.KITTI360FisheyeDataset", raw_path=raw, split_file=meta, frame_ids=[0, 1, -1],
                 is_filter_static=False, use_right_image=True, fisheye_mask=mask_path,
                 augmentation=edict(name="fsnet_b200.data.device_augment.DeviceAugmentation", pipeline=fisheye_train_cfg()))
    check([fish[i] for i in range(4)], fish.transform, exact_original=False)
