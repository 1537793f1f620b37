"""Host-side augmentations (fsnet_b200/data/augmentations.py, exposed under the reference's dotted names) against golden
vectors produced by the reference's own classes and builder (tests/golden/make_golden_aug.py): the train and the
validation lists of configs/kitti_wpose_example, three consecutive seeded samples each."""
import os

import numpy as np
import pytest
import torch

from aug_cases import raw_sample, train_cfg, val_cfg, summarize


@pytest.mark.parametrize("name,cfg_fn,seed", [("aug_train", train_cfg, 7), ("aug_val", val_cfg, 8)])
def test_pipeline_matches_reference(golden_dir, name, cfg_fn, seed):
    from vision_base.utils.builder import build
    import vision_base
    assert "reference" not in vision_base.__path__[0]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    np.random.seed(seed)
    pipe = build(**cfg_fn())
    for i in range(3):
        out = pipe(raw_sample(100 + i))
        mine = summarize(out)
        keys = [k[len(f"{i}/"):] for k in g.files if k.startswith(f"{i}/")]
        assert sorted(keys) == sorted(mine.keys())
        for k in keys:
            want, got = g[f"{i}/{k}"], mine[k]
            if k.startswith("dtype/"):
                assert str(want) == str(got), k            # incl. the fp64 patched_mask
            elif k.startswith("full/relative_pose"):
                # the reference goes through Euler angles, this repo conjugates with the reflection: same pose to rounding
                np.testing.assert_allclose(got, want, atol=2e-6)
            else:
                np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6, err_msg=k)
        assert isinstance(out[("image", 0)], torch.Tensor) and out[("image", 0)].shape == (3, 48, 160)


def test_flip_relative_pose_is_an_involution_and_mirrors_translation():
    from vision_base.data.augmentations.utils import flip_relative_pose
    T = raw_sample(3)[("relative_pose", 1)]
    F = flip_relative_pose(T.copy(), 0)
    np.testing.assert_allclose(flip_relative_pose(F.copy(), 0), T, atol=1e-6)
    assert np.isclose(F[0, 3], -T[0, 3]) and np.allclose(F[1:3, 3], T[1:3, 3])
    np.testing.assert_allclose(F[:3, :3] @ F[:3, :3].T, np.eye(3), atol=1e-5)


def test_depth_metrics_match_reference(golden_dir):
    """compute_errors / compute_depth_errors / disp<->depth and the Eigen crop + median-scaling protocol against values
    computed by the reference's own functions (tests/golden/make_golden_aug.py::run_metrics)."""
    from monodepth.networks.utils import monodepth_utils as U
    from fsnet_b200.utils.metrics import eigen_median_scaled_errors
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    r = np.random.default_rng(21)
    gt = r.uniform(1.0, 70.0, size=5000)
    pred = gt * r.uniform(0.7, 1.4, size=5000)
    np.testing.assert_allclose(np.array(U.compute_errors(gt, pred)), g["errors"], rtol=1e-12)
    np.testing.assert_allclose([float(v) for v in U.compute_depth_errors(torch.tensor(gt), torch.tensor(pred))], g["torch_errors"], rtol=1e-10)
    np.testing.assert_allclose(U.disp_to_depth(np.linspace(0, 1, 11), 0.1, 100.0)[1], g["depth"], rtol=1e-12)
    np.testing.assert_allclose(U.depth_to_disp(np.linspace(0.5, 90, 11), 0.1, 100.0), g["disp"], rtol=1e-12)
    gt_map = np.where(r.uniform(size=(94, 310)) < 0.2, r.uniform(0.5, 90.0, size=(94, 310)), 0.0).astype(np.float32)
    pred_map = r.uniform(2.0, 60.0, size=(48, 160)).astype(np.float32)
    res = eigen_median_scaled_errors(pred_map, gt_map)
    np.testing.assert_allclose(res["ratio"], g["eigen_ratio"], rtol=1e-6)
    np.testing.assert_allclose(np.array(res["error"]), g["eigen_error"], rtol=1e-5)
    np.testing.assert_allclose(np.array(res["abs_error"]), g["eigen_abs_error"], rtol=1e-5)
