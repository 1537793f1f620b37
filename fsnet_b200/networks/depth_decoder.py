"""U-Net depth decoders (reference: monodepth/networks/models/heads/depth_encoder.py:17-139).

``DepthDecoder`` = sigmoid disparity head; ``MultiChannelDepthDecoder`` = softmax over log-spaced depth
bins (the one every shipped config uses).  Same constructor, ``decoder`` ModuleList order and
``depth_bins`` buffer as the reference, so checkpoints interchange.  The head arithmetic
(clamp/softmax/expectation/depth->disp) is the fused CUDA kernel ``fsnet_depth_head_*``."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .blocks import ConvBnReLU
from .. import functional as Fn


class DepthDecoder(nn.Module):
    multi_channel = False

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True, min_depth=0.1, max_depth=100,
                 base_fx=None):
        super().__init__()
        self.num_output_channels = num_output_channels
        self.use_skips = use_skips
        self.upsample_mode = "nearest"
        self.scales = scales
        self.base_fx = base_fx
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = np.array([16, 32, 64, 128, 256])
        self.min_depth, self.max_depth = min_depth, max_depth
        lo, hi = np.log(min_depth), np.log(max_depth)          # depth_encoder.py:68-74
        self.register_buffer("depth_bins", torch.exp(torch.arange(lo, hi, (hi - lo) / num_output_channels)))
        self.convs = OrderedDict()
        for i in range(4, -1, -1):
            cin = int(num_ch_enc[-1] if i == 4 else self.num_ch_dec[i + 1])
            cout = int(self.num_ch_dec[i])
            self.convs[("upconv", i, 0)] = ConvBnReLU(cin, cout, kernel_size=(3, 3))
            cin = cout + (int(num_ch_enc[i - 1]) if (use_skips and i > 0) else 0)
            self.convs[("upconv", i, 1)] = ConvBnReLU(cin, cout, kernel_size=(3, 3), padding_mode="replicate")
        for s in self.scales:
            self.convs[("dispconv", s)] = nn.Conv2d(int(self.num_ch_dec[s]), num_output_channels, kernel_size=3, padding=1,
                                                   padding_mode="replicate")
        self.decoder = nn.ModuleList(list(self.convs.values()))
        self.sigmoid = nn.Sigmoid()

    def _get_scale(self, P2):
        """fx / base_fx per sample, or None (depth_encoder.py:36-43)."""
        if self.base_fx is None or P2 is None:
            return None
        return (P2[:, 0, 0] / self.base_fx).float()

    def _trunk(self, input_features):
        """Encoder + decoder as one autograd node of the tcgen05 executor.  ``input_features`` is the deferred handle ResNet.forward
        returned; if some other code has already indexed it (the list then holds exported copies of the five feature maps), the
        network is still executed from the image, so gradients and BatchNorm statistics are those of the real path."""
        from .ops_tc import LazyFeatures, runner_for
        if isinstance(input_features, LazyFeatures):
            logits = runner_for(self, input_features.backbone).depth_logits(input_features.image)
            for i in range(4, -1, -1):
                if i in self.scales:
                    yield i, logits[i]
            return
        if ops.COMPARATOR is None:
            raise TypeError("DepthDecoder expects the deferred features returned by fsnet_b200's ResNet.forward. " + ops.NO_CPU)
        yield from ops.COMPARATOR.decoder_trunk(self, input_features)

    def forward(self, input_features, P2=None):
        outputs = {}
        scale = self._get_scale(P2)
        for i, logits in self._trunk(input_features):
            outputs[("logits", i)] = logits
            depth, disp = Fn.depth_head(logits, None, scale, True, self.min_depth, self.max_depth)
            outputs[("disp", i)] = disp
            outputs[("depth", i, i)] = depth
        return outputs


class MultiChannelDepthDecoder(DepthDecoder):
    multi_channel = True

    def forward(self, input_features, P2=None):
        outputs = {}
        scale = self._get_scale(P2)
        for i, logits in self._trunk(input_features):
            outputs[("logits", i)] = logits
            outputs[("depth", i, i)], outputs[("disp", i)] = Fn.depth_head(
                logits, self.depth_bins, scale, False, self.min_depth, self.max_depth)
        return outputs


class MultiChannelDepthDecoderUncertain(DepthDecoder):
    """Second-stage (distillation) decoder: the softmax-bins depth head plus a one-channel uncertainty head per scale
    (depth_encoder.py:142-194).  ``('uncertain_logit', s)`` is what the fused distillation loss consumes;
    ``('uncertain_z', s)`` = sigmoid of it is the reference's output entry (detached: informational)."""
    multi_channel = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for s in self.scales:
            self.convs[("uncertain_logz", s)] = nn.Conv2d(int(self.num_ch_dec[s]), 1, kernel_size=3, padding=1, padding_mode="replicate")
        self.decoder = nn.ModuleList(list(self.convs.values()))       # same ModuleList order as the reference: heads last

    def _trunk_with_uncertainty(self, input_features):
        from .ops_tc import LazyFeatures, runner_for
        if isinstance(input_features, LazyFeatures):
            outs = runner_for(self, input_features.backbone).depth_logits(input_features.image)
            for i in range(4, -1, -1):
                if i in self.scales:
                    yield i, outs[i], outs[("uncertain", i)]
            return
        if ops.COMPARATOR is None:
            raise TypeError("DepthDecoder expects the deferred features returned by fsnet_b200's ResNet.forward. " + ops.NO_CPU)
        yield from ops.COMPARATOR.decoder_trunk(self, input_features, with_uncertainty=True)

    def forward(self, input_features, P2=None):
        outputs = {}
        scale = self._get_scale(P2)
        for i, logits, ulogit in self._trunk_with_uncertainty(input_features):
            outputs[("logits", i)] = logits
            outputs[("depth", i, i)], outputs[("disp", i)] = Fn.depth_head(
                logits, self.depth_bins, scale, False, self.min_depth, self.max_depth)
            outputs[("uncertain_logit", i)] = ulogit
            outputs[("uncertain_z", i)] = torch.sigmoid(ulogit.detach())
        return outputs
