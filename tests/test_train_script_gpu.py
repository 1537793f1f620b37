"""scripts/train.py (reference CLI, scripts/train.py:21-214) end to end on the three synthetic configs: builder ->
dataloader -> BaseTrainingHook -> FusedAdam -> scheduler -> checkpoint, a few steps each."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config,extra", [
    ("kitti_wpose_synthetic.py", ["--data.batch_size=2"]),
    ("kitti_posenet_synthetic.py", ["--data.batch_size=2"]),
    ("kitti360_fisheye_synthetic.py", ["--data.batch_size=2"]),
])
def test_train_script_runs(tmp_path, config, extra):
    env = dict(os.environ, FSNET_WORKDIR=str(tmp_path), PYTHONPATH=REPO)
    cmd = [sys.executable, os.path.join(REPO, "scripts", "train.py"), f"--config={os.path.join(REPO, 'configs', config)}",
           "--experiment_name=pytest", "--trainer.max_steps=4", "--trainer.max_epochs=1", "--data.num_workers=0",
           "--train_dataset.length=16"] + extra
    out = subprocess.run(cmd, env=env, cwd=REPO, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "finished 4 steps" in out.stdout
    ckpts = [f for _, _, fs in os.walk(tmp_path) for f in fs if f.endswith("_latest.pth")]
    assert ckpts, "no checkpoint written"
    state = torch.load([os.path.join(d, f) for d, _, fs in os.walk(tmp_path) for f in fs if f.endswith("_latest.pth")][0],
                       map_location="cpu")
    assert "model_state_dict" in state and "optimizer_state_dict" in state
    assert all(torch.isfinite(v).all() for v in state["model_state_dict"].values() if v.is_floating_point())


def test_train_script_on_kitti_files(tmp_path):
    """The KITTI recipe end to end on FILES: miniature KITTI tree -> reference-named readers -> augmentation lists ->
    dataloader workers -> tcgen05 / fused-loss training step -> checkpoint."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from kitti_fixture import build_tree
    raw, split = build_tree(str(tmp_path / "kitti"))
    env = dict(os.environ, FSNET_WORKDIR=str(tmp_path), PYTHONPATH=REPO, FSNET_KITTI_PATH=raw, FSNET_KITTI_SPLIT=split,
               FSNET_SHIFT_BORDER="32")
    cmd = [sys.executable, os.path.join(REPO, "scripts", "train.py"), f"--config={os.path.join(REPO, 'configs', 'kitti_wpose_files.py')}",
           "--experiment_name=pytest", "--trainer.max_steps=3", "--trainer.max_epochs=2", "--data.batch_size=2", "--data.num_workers=2"]
    out = subprocess.run(cmd, env=env, cwd=REPO, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "finished 3 steps" in out.stdout


def test_train_script_on_kitti360_fisheye_files(tmp_path):
    """The KITTI-360 fisheye recipe end to end on FILES: MEI yaml calibration -> calib_meta -> device-built ray table -> FishEyeDecoder."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from kitti_fixture import build_kitti360_tree
    raw, meta, mask = build_kitti360_tree(str(tmp_path / "k360"))
    env = dict(os.environ, FSNET_WORKDIR=str(tmp_path), PYTHONPATH=REPO, FSNET_KITTI360_PATH=raw, FSNET_KITTI360_SPLIT=meta,
               FSNET_FISHEYE_MASK=mask, FSNET_FISHEYE_SIZE="64")
    cmd = [sys.executable, os.path.join(REPO, "scripts", "train.py"), f"--config={os.path.join(REPO, 'configs', 'kitti360_fisheye_files.py')}",
           "--experiment_name=pytest", "--trainer.max_steps=3", "--trainer.max_epochs=3", "--data.batch_size=2", "--data.num_workers=0"]
    out = subprocess.run(cmd, env=env, cwd=REPO, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "finished 3 steps" in out.stdout
