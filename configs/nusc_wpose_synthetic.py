"""cfg4 (BASELINE.json configs[3]): nuScenes 640x320, ResNet-18 depth net with dataset poses (the meta-architecture of the
reference's configs/nusc_wpose_example:178-212), batch 8 per GPU, on the synthetic triplet dataset.  The shipped nuScenes recipe
itself (ResNet-34, 64 bins, base_fx, 288x512, file readers) is configs/nusc_wpose_files.py."""
import os

from vision_base.utils.utils import cfg_from_file

cfg = cfg_from_file(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kitti_wpose_synthetic.py"))
H, W = 320, 640
cfg.data.batch_size = 8
cfg.data.rgb_shape = (H, W, 3)
for ds in (cfg.train_dataset, cfg.val_dataset):
    ds.height, ds.width = H, W
cfg.meta_arch.head_cfg.height, cfg.meta_arch.head_cfg.width = H, W
