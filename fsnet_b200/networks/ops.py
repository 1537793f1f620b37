"""Composite layer operations used by the network modules.

Two back-ends:
  "tc"     hand-written tcgen05 implicit-GEMM convolutions + fused BN statistics (fsnet_b200/csrc/conv_tc.cu)
  "torch"  stock PyTorch ops (cuDNN) -- the INTERIM library path for layer types the tcgen05
           path does not cover yet; DESIGN.md lists which.
The network modules only hold parameters (reference names / state-dict layout) and call these.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

BACKEND = "torch"


def set_backend(name: str) -> None:
    global BACKEND
    assert name in ("torch", "tc")
    BACKEND = name


def tc_available() -> bool:
    """True when the tcgen05 convolution kernels are built into the library."""
    try:
        from . import ops_tc  # noqa: F401
        return ops_tc.available()
    except ImportError:
        return False


def precision_note() -> str:
    if BACKEND == "tc":
        return "tcgen05 bf16x3 split operands, fp32 accumulate"
    return "cuDNN fp32 (interim library path)"


def conv_bn_act(x, conv: nn.Conv2d, bn, relu: bool = True, residual=None):
    """conv -> (train-mode) batch-norm -> (+ residual) -> ReLU."""
    if BACKEND == "tc" and x.is_cuda:
        from . import ops_tc
        if ops_tc.supports(conv, x):
            return ops_tc.conv_bn_act(x, conv, bn, relu, residual)
    y = conv(x)
    if bn is not None:
        y = bn(y)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def conv_act(x, conv: nn.Conv2d, relu: bool = False):
    if BACKEND == "tc" and x.is_cuda:
        from . import ops_tc
        if ops_tc.supports(conv, x):
            return ops_tc.conv_bn_act(x, conv, None, relu, None)
    y = conv(x)
    return F.relu(y) if relu else y


def maxpool3x3s2(x):
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


def upsample2x_concat(x, skip=None):
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    return x if skip is None else torch.cat([x, skip], 1)
