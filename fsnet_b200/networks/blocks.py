"""ConvBnReLU (reference: vision_base/networks/blocks/blocks.py:33-54).

Parameter container with the reference's state-dict layout (``sequence.0`` = conv WITH bias,
``sequence.1`` = BatchNorm2d).  ReLU is always applied -- the reference ignores its ``relu`` argument
(blocks.py:47)."""
import torch.nn as nn

from . import ops


class ConvBnReLU(nn.Module):
    def __init__(self, input_features=1, output_features=1, kernel_size=(1, 1), stride=[1, 1], padding="SAME", dilation=1,
                 groups=1, relu=True, **kwargs):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size, kernel_size)
        pad = int((kernel_size[0] - 1) / 2) * dilation if padding.lower() == "same" else 0
        self.sequence = nn.Sequential(
            nn.Conv2d(input_features, output_features, kernel_size=kernel_size, stride=stride, padding=pad,
                      dilation=dilation, groups=groups, **kwargs),
            nn.BatchNorm2d(output_features))
        self.relu = True

    def forward(self, x):
        if ops.COMPARATOR is None:
            raise RuntimeError("ConvBnReLU is a parameter container on the tcgen05 path. " + ops.NO_CPU)
        return ops.COMPARATOR.conv_bn_relu_forward(self, x)
