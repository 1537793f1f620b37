"""cfg3 (BASELINE.json configs[2]): KITTI-360 pinhole 768x192, ResNet-50, 4 scales, batch 8 per GPU.  Topology of the reference's
configs/multi_dataset_example:225-262 (ResNet-50 encoder, num_ch_enc = 64/256/512/1024/2048) at the BASELINE size, on the synthetic
triplet dataset: kitti_wpose_synthetic.py with the entries below replaced."""
import os

import numpy as np

from vision_base.utils.utils import cfg_from_file

cfg = cfg_from_file(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kitti_wpose_synthetic.py"))
H, W = 192, 768
cfg.data.batch_size = 8
cfg.data.rgb_shape = (H, W, 3)
for ds in (cfg.train_dataset, cfg.val_dataset):
    ds.height, ds.width = H, W
cfg.meta_arch.depth_backbone_cfg.depth = 50
cfg.meta_arch.head_cfg.height, cfg.meta_arch.head_cfg.width = H, W
cfg.meta_arch.head_cfg.depth_decoder_cfg.num_ch_enc = np.array([64, 256, 512, 1024, 2048])
