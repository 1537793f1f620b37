"""Composite layer operations used by the network modules.

Two back-ends:
  "tc"     (default) the whole network runs as hand-written tcgen05 kernels through fsnet_b200/engine.py:
           ``ResNet.forward`` only returns a deferred handle and the consuming head executes encoder + decoder
           as one autograd node.  The functions below are then not on the path at all.
  "torch"  stock PyTorch ops (cuDNN): a comparison / debugging path and the route for the few constructor
           options the tcgen05 executor rejects (norm_eval=True, frozen_stages, dilation), never a silent
           fallback: the executor raises NotImplementedError and the user selects this back-end explicitly
           (``ops.set_backend("torch")`` or FSNET_CONV_BACKEND=torch).
The network modules only hold parameters (reference names / state-dict layout).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

BACKEND = os.environ.get("FSNET_CONV_BACKEND", "tc")      # "tc" needs CUDA tensors; CPU tensors always take the torch ops


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
    return "cuDNN fp32 (comparison path)"


def conv_bn_act(x, conv: nn.Conv2d, bn, relu: bool = True, residual=None):
    """conv -> (train-mode) batch-norm -> (+ residual) -> ReLU."""
    y = conv(x)
    if bn is not None:
        y = bn(y)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def conv_act(x, conv: nn.Conv2d, relu: bool = False):
    y = conv(x)
    return F.relu(y) if relu else y


def maxpool3x3s2(x):
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


def upsample2x_concat(x, skip=None):
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    return x if skip is None else torch.cat([x, skip], 1)
