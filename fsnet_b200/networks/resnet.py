"""ResNet-18/34/50/101/152 encoder returning five feature maps.

Mirror of vision_base/networks/models/backbone/resnet.py (constructor arguments :96-105, forward
:199-213, stage/BN freezing :169-197, 6-channel stem for the PoseNet :119,155-160) with the same
attribute names, hence the same state-dict keys.  The modules only hold parameters: the arithmetic is the tcgen05 executor
(fsnet_b200/engine.py), see networks/ops.py."""
import math
from typing import Tuple

import torch
import torch.nn as nn

from . import ops

_LAYERS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        if ops.COMPARATOR is None:
            raise RuntimeError("BasicBlock is a parameter container on the tcgen05 path. " + ops.NO_CPU)
        return ops.COMPARATOR.block_forward(self, x)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        if ops.COMPARATOR is None:
            raise RuntimeError("Bottleneck is a parameter container on the tcgen05 path. " + ops.NO_CPU)
        return ops.COMPARATOR.block_forward(self, x)


class ResNet(nn.Module):
    planes = [64, 128, 256, 512]

    def __init__(self, block, layers: Tuple[int, ...], num_stages: int = 4, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1),
                 out_indices=(-1, 0, 1, 2, 3), frozen_stages: int = -1, norm_eval: bool = True, num_input_images=1):
        self.inplanes = 64
        super().__init__()
        assert 1 <= num_stages <= 4 and max(out_indices) < num_stages
        self.num_stages, self.strides, self.dilations = num_stages, strides, dilations
        self.out_indices, self.frozen_stages = out_indices, frozen_stages
        self.num_input_images, self.norm_eval = num_input_images, norm_eval
        self.conv1 = nn.Conv2d(3 * num_input_images, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        for i in range(num_stages):
            setattr(self, f"layer{i + 1}", self._make_layer(block, self.planes[i], layers[i], strides[i], dilations[i]))
        for m in self.modules():          # fan-out kaiming normal / BN (1, 0): resnet.py:126-132
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self.train()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                                       nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes, dilation=dilation) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def load_state_dict(self, state_dict, *args, **kwargs):
        # ImageNet 3-channel stem -> N-image stem: tile and rescale (resnet.py:155-160)
        if "conv1.weight" in state_dict and self.conv1.weight.shape != state_dict["conv1.weight"].shape:
            state_dict["conv1.weight"] = torch.cat([state_dict["conv1.weight"]] * self.num_input_images, 1) / self.num_input_images
        return super().load_state_dict(state_dict, *args, **kwargs)

    def train(self, mode=True):
        super().train(mode)
        if mode:
            self._freeze_stages()
            if self.norm_eval:
                for m in self.modules():
                    if isinstance(m, nn.modules.batchnorm._BatchNorm):
                        m.eval()
        return self

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            for m in (self.conv1, self.bn1):
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f"layer{i}")
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def forward(self, img_batch):
        if ops.COMPARATOR is not None:
            return ops.COMPARATOR.resnet_forward(self, img_batch)
        if not img_batch.is_cuda:
            raise RuntimeError(ops.NO_CPU)
        from .ops_tc import LazyFeatures
        return LazyFeatures(self, img_batch)      # executed by the consuming head (fsnet_b200/engine.py)


def resnet(depth, pretrained=True, **kwargs):
    """Factory with the reference's signature (resnet.py:270-284).  ``pretrained=True`` needs the
    torchvision ImageNet checkpoint, which the reference downloads (resnet.py:224); offline it must be
    given as a local file through the FSNET_PRETRAINED_DIR environment variable."""
    if depth not in _LAYERS:
        raise ValueError("Unsupported model depth, must be one of 18, 34, 50, 101, 152")
    model = ResNet(BasicBlock if depth < 50 else Bottleneck, _LAYERS[depth], **kwargs)
    if pretrained:
        import os
        d = os.environ.get("FSNET_PRETRAINED_DIR")
        path = None if d is None else os.path.join(d, f"resnet{depth}.pth")
        if path is None or not os.path.exists(path):
            raise FileNotFoundError(
                f"pretrained=True needs resnet{depth}.pth in $FSNET_PRETRAINED_DIR (no network access to download it); "
                "use pretrained=False for synthetic runs")
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
    return model
