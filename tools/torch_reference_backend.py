"""Stock-PyTorch (cuDNN) execution of the fsnet_b200 network modules: the COMPARATOR, not a product path.

The product modules (fsnet_b200/networks) only hold parameters; their arithmetic is the tcgen05 executor (fsnet_b200/engine.py)
and they raise on CPU tensors.  ``enable()`` plugs this file's forward functions into them so that the very same module tree
(same parameters, same state dict) runs through ``torch.nn.functional`` -- used by ``bench.py --backend torch`` (the informational
"what does cuDNN do on this box" line) and by tests that compare feature maps.  ``disable()`` restores the product behaviour."""
import torch
import torch.nn.functional as F


def conv_bn_act(x, conv, bn, relu=True, residual=None):
    y = conv(x)
    if bn is not None:
        y = bn(y)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def conv_act(x, conv, relu=False):
    y = conv(x)
    return F.relu(y) if relu else y


def block_forward(block, x):
    """BasicBlock / Bottleneck (reference resnet.py:34-50, 71-89)."""
    res = x if block.downsample is None else conv_bn_act(x, block.downsample[0], block.downsample[1], relu=False)
    out = conv_bn_act(x, block.conv1, block.bn1, relu=True)
    if hasattr(block, "conv3"):
        out = conv_bn_act(out, block.conv2, block.bn2, relu=True)
        return conv_bn_act(out, block.conv3, block.bn3, relu=True, residual=res)
    return conv_bn_act(out, block.conv2, block.bn2, relu=True, residual=res)


def conv_bn_relu_forward(m, x):
    return conv_bn_act(x, m.sequence[0], m.sequence[1], relu=True)


def resnet_forward(net, img):
    outs = []
    x = conv_bn_act(img, net.conv1, net.bn1, relu=True)
    if -1 in net.out_indices:
        outs.append(x)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for i in range(net.num_stages):
        for block in getattr(net, f"layer{i + 1}"):
            x = block_forward(block, x)
        if i in net.out_indices:
            outs.append(x)
    return outs


def decoder_trunk(dec, feats, with_uncertainty=False):
    x = feats[-1]
    for i in range(4, -1, -1):
        x = conv_bn_relu_forward(dec.convs[("upconv", i, 0)], x)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if dec.use_skips and i > 0:
            x = torch.cat([x, feats[i - 1]], 1)
        x = conv_bn_relu_forward(dec.convs[("upconv", i, 1)], x)
        if i in dec.scales:
            logits = conv_act(x, dec.convs[("dispconv", i)])
            if with_uncertainty:
                yield i, logits, conv_act(x, dec.convs[("uncertain_logz", i)])
            else:
                yield i, logits


def pose_decoder_forward(dec, input_features):
    last = [f[-1] for f in input_features]
    cat = torch.cat([conv_act(f, dec.convs["squeeze"], relu=True) for f in last], 1)
    out = conv_act(cat, dec.convs[("pose", 0)], relu=True)
    out = conv_act(out, dec.convs[("pose", 1)], relu=True)
    out = conv_act(out, dec.convs[("pose", 2)], relu=False)
    out = out.float().mean(3).mean(2)
    out = 0.01 * out.view(-1, dec.num_frames_to_predict_for, 1, 6)
    return out[..., :3], out[..., 3:]


def enable():
    import sys
    from fsnet_b200.networks import ops
    ops.COMPARATOR = sys.modules[__name__]


def disable():
    from fsnet_b200.networks import ops
    ops.COMPARATOR = None
